"""HBM-resident embedding store: what replaces the reference's python dict
(app_pipeline.py:18) and its Milvus collection (ImageEmbeddingSystem.py:41-61).

Layout: one contiguous row-major (N, D) matrix on the device (fp32, or bf16 for the tcgen05
path), a host list of paths in insertion order (row i <-> paths[i]) and optional fp32 magnitudes.
"""
import numpy as np
import torch

from . import ops


class EmbeddingStore:
    def __init__(self, dim=None, dtype=torch.float32):
        self.dim = dim
        self.dtype = dtype
        self.paths = []
        self.matrix = None          # (N, D) CUDA tensor
        self.magnitudes = None      # (N,) CUDA fp32 or None
        self._pending = []          # host rows not yet uploaded
        self._pending_mag = []
        self.version = 0            # bumped whenever rows / paths change: keys the caches built on top of the store

    def __len__(self):
        return len(self.paths)

    def add(self, path, vector, magnitude=None):
        v = np.asarray(vector, dtype=np.float32).reshape(-1)
        if self.dim is None:
            self.dim = v.shape[0]
        if v.shape[0] != self.dim:
            raise ValueError(f"embedding dimension {v.shape[0]} != store dimension {self.dim}")
        self.paths.append(str(path))
        self.version += 1
        self._pending.append(v)
        self._pending_mag.append(1.0 if magnitude is None else float(magnitude))

    def add_batch(self, paths, matrix, magnitudes=None):
        """Append rows that may already live on the device (no host round trip)."""
        self.flush()
        m = ops.as_device_matrix(matrix, dtype=self.dtype)
        if self.dim is None:
            self.dim = m.shape[1]
        if m.shape[1] != self.dim or len(paths) != m.shape[0]:
            raise ValueError("add_batch: shape mismatch")
        mag = torch.ones(m.shape[0], dtype=torch.float32, device=m.device) if magnitudes is None else \
            torch.as_tensor(magnitudes, dtype=torch.float32).to(m.device)
        self.paths.extend(str(p) for p in paths)
        self.version += 1
        self.matrix = m if self.matrix is None else torch.cat([self.matrix, m])
        self.magnitudes = mag if self.magnitudes is None else torch.cat([self.magnitudes, mag])

    def flush(self):
        if self._pending:
            rows, mags = np.stack(self._pending), np.asarray(self._pending_mag, dtype=np.float32)
            self._pending, self._pending_mag = [], []
            m = ops.as_device_matrix(rows, dtype=self.dtype)
            g = torch.from_numpy(mags).to(m.device)
            self.matrix = m if self.matrix is None else torch.cat([self.matrix, m])
            self.magnitudes = g if self.magnitudes is None else torch.cat([self.magnitudes, g])
        return self.matrix

    def device_matrix(self):
        return self.flush()

    def prepared(self):
        """The matrix with its per-store search state (ops.PreparedIndex: row norms, bf16 planes of an fp32 store), built
        once per content version - what a query BATCH against a static store should be searched with."""
        m = self.flush()
        if m is None:
            return None
        key = (self.version, m.data_ptr(), tuple(m.shape))
        if getattr(self, "_prepared_key", None) != key:
            self._prepared = ops.prepare_index(m)
            self._prepared_key = key
        return self._prepared

    def set_path(self, row, path):
        """Rename a stored row (keeps the de-duplication groups of the searchers in step)."""
        self.paths[row] = str(path)
        self.version += 1
