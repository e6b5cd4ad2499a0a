"""ImageEmbeddingSystem - the reference's embedding producer / store surface
(/root/reference/src/ImageEmbeddingSystem.py:18-215) on the B200 hot path.

Kept: generate_embedding -> (unit vector, magnitude), process_and_store_images -> (ok, failed),
get_embeddings, get_embeddings_with_magnitude, reconstruct_original_embeddings, and the
skip-and-count error behaviour.  Replaced: the CLIP image tower by the 512-bin colour-histogram
kernel north_star names (EMBEDDING_DIM = 8*8*8), and the Milvus collection by an HBM-resident
matrix (store.EmbeddingStore).  `model` / `processor` are accepted for signature compatibility
and ignored.  `image_size=224` turns on the processor's image front-end (:82-83: shorter edge -> 224
with PIL BICUBIC, centre crop 224 x 224) on the device, bit-exact with PIL (ops.resize_crop); the
default None embeds the image at its own size.
"""
import logging
from pathlib import Path
from typing import List, Tuple

import numpy as np

from . import ops
from .config import EMBEDDING_DIM
from .store import EmbeddingStore

logger = logging.getLogger(__name__)


class ImageEmbeddingSystem:
    """Handles image embedding generation and storage (device-resident)."""

    def __init__(self, model=None, processor=None, device: str = "cuda", colorspace: str = "rgb", image_size=None):
        self.model = model
        self.processor = processor
        self.device = device
        self.colorspace = colorspace
        self.image_size = image_size
        self.setup_milvus()

    def setup_milvus(self):
        """Reference: connect to Milvus and create the collection (:35-66).  Here: create the
        HBM-resident store (paths, unit vectors, magnitudes)."""
        self.collection = EmbeddingStore(dim=EMBEDDING_DIM)

    @staticmethod
    def _load(image):
        if isinstance(image, np.ndarray):
            arr = image
        else:
            from PIL import Image
            with Image.open(image) as im:
                arr = np.asarray(im.convert("RGB"), dtype=np.uint8)
        if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[2] != 3:
            raise ValueError(f"expected an RGB uint8 image, got {arr.dtype} {arr.shape}")
        return np.ascontiguousarray(arr)

    def embed_batch(self, images):
        """(B,H,W,3) uint8 -> (unit (B,512) fp32, magnitude (B,) fp32) device tensors."""
        if self.image_size is not None and tuple(images.shape[-3:-1]) != (self.image_size, self.image_size):
            images = ops.resize_crop(images, self.image_size)
        counts = ops.histogram(images, self.colorspace)
        _raw, unit, mag = ops.counts_to_embedding(counts)
        return unit, mag

    def generate_embedding(self, image_path) -> Tuple[np.ndarray, float]:
        """Embedding of one image: (normalized_embedding, magnitude) (:68-98).  Raises on failure."""
        try:
            arr = self._load(image_path)
            unit, mag = self.embed_batch(arr[None])
            return unit[0].cpu().numpy(), float(mag[0].item())
        except Exception as e:
            logger.error(f"Failed to generate embedding for {image_path}: {e}")
            raise

    def process_and_store_images(self, image_paths: List[Path]) -> Tuple[int, int]:
        """Embed and store images; returns (successful_count, failed_count) (:100-145)."""
        if not image_paths:
            logger.warning("No image paths provided for processing.")
            return 0, 0
        failed_count = 0
        paths, arrays = [], []
        for image_path in image_paths:
            try:
                arrays.append(self._load(image_path))
                paths.append(str(image_path))
            except Exception as e:  # noqa: BLE001 - reference skips and counts (:126-129)
                logger.warning(f"Skipping {image_path} due to error: {e}")
                failed_count += 1
        # same-shape images go through the kernel as one batch
        groups = {}
        for p, a in zip(paths, arrays):
            groups.setdefault(a.shape, []).append((p, a))
        for shape, items in groups.items():
            unit, mag = self.embed_batch(np.stack([a for _, a in items]))
            self.collection.add_batch([p for p, _ in items], unit, mag)
        return len(paths), failed_count

    def store_arrays(self, paths, images):
        """Device fast path: embed a (B,H,W,3) uint8 batch (host or device) and append it."""
        unit, mag = self.embed_batch(images)
        self.collection.add_batch(paths, unit, mag)
        return len(paths)

    def get_embeddings(self, limit: int = 1000) -> List[Tuple[str, np.ndarray]]:
        """(image_path, normalized_embedding) tuples, at most `limit` (:147-171)."""
        m = self.collection.device_matrix()
        if m is None:
            return []
        host = m[:limit].float().cpu().numpy()
        return [(p, host[i]) for i, p in enumerate(self.collection.paths[:limit])]

    def get_embeddings_with_magnitude(self, limit: int = 1000) -> List[Tuple[str, np.ndarray, float]]:
        """(image_path, normalized_embedding, magnitude) tuples (:173-202)."""
        m = self.collection.device_matrix()
        if m is None:
            return []
        host = m[:limit].float().cpu().numpy()
        mags = self.collection.magnitudes[:limit].cpu().numpy()
        return [(p, host[i], float(mags[i])) for i, p in enumerate(self.collection.paths[:limit])]

    def reconstruct_original_embeddings(self, embeddings):
        """Unnormalised embeddings from (path, unit, magnitude) triples (:204-215)."""
        return [(path, emb * mag) for path, emb, mag in embeddings]
