"""Oracle (test infrastructure): seeded synthetic inputs (SURVEY.md section 8d shapes)."""
import numpy as np


def gaussian(n, d, seed, normalize=False):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    if normalize:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def relu_features(n, d, seed):
    """Non-negative, ~50 % sparse ResNet-like features (config 3)."""
    return np.maximum(gaussian(n, d, seed), 0.0)


def images_uniform(b, h, w, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(b, h, w, 3), dtype=np.uint8)


def images_palette(b, h, w, seed):
    """3-5 base colours, random axis-aligned rectangles, +-8 noise: peaky histograms, many ties."""
    rng = np.random.default_rng(seed)
    out = np.empty((b, h, w, 3), dtype=np.uint8)
    for i in range(b):
        ncol = int(rng.integers(3, 6))
        pal = rng.integers(0, 256, size=(ncol, 3))
        img = np.empty((h, w, 3), dtype=np.int64)
        img[:] = pal[0]
        for _ in range(int(rng.integers(3, 9))):
            y0, y1 = np.sort(rng.integers(0, h + 1, size=2))
            x0, x1 = np.sort(rng.integers(0, w + 1, size=2))
            img[y0:y1, x0:x1] = pal[int(rng.integers(0, ncol))]
        img += rng.integers(-8, 9, size=img.shape)
        out[i] = np.clip(img, 0, 255).astype(np.uint8)
    return out
