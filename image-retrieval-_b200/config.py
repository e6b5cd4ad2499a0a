"""Constants of the reference's config.py (/root/reference/src/config.py:3-7)."""

MILVUS_HOST = 'localhost'      # kept for API completeness; the vector store is an HBM-resident matrix
MILVUS_PORT = '19530'
BATCH_SIZE = 100
EMBEDDING_DIM = 512            # == 8*8*8 histogram bins
SCORE_THRESHOLD = 0.25
