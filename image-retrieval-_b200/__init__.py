"""B200-native brute-force retrieval hot path (distances + fused top-k + colour histograms).

Drop-in for the reference's Python call surface on this path:

    from image_retrieval_b200.geometric_metrics import GeometricSimilarityMetrics
    from image_retrieval_b200.app_pipeline import EnhancedImageSearchApp, SimpleSearcher
    from image_retrieval_b200.image_search import EnhancedTextImageSearcher
    from image_retrieval_b200.ImageEmbeddingSystem import ImageEmbeddingSystem
    from image_retrieval_b200 import ops          # batched API over the C ABI (include/b200ir.h)

All arithmetic runs in hand-written sm_100a CUDA behind libb200ir.so; there is no CPU fallback:
calling a compute entry point without the library or without a B200 raises.
"""
__version__ = "0.1.0"
