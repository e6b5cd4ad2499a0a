"""Oracle (test infrastructure): the image front-end either side of the embedding producer.

The reference opens an image with PIL and hands it to CLIPProcessor (/root/reference/src/ImageEmbeddingSystem.py:82-83,
app_pipeline.py:103-104,127-131), whose image pipeline is: resize the SHORTER edge to 224 with PIL BICUBIC
(transformers image_transforms.get_resize_output_image_size: new_long = int(224 * long / short)), centre-crop
224 x 224 (top = (h - 224) // 2, left = (w - 224) // 2), then float rescale / normalise for the CLIP tower (out of
scope: the build's embedding is the colour histogram of the cropped uint8 image).

The resampling arithmetic is third-party and absent from /root/reference: Pillow (unpinned in requirements.txt; 12.2.0
in this image), src/libImaging/Resample.c.  Its published 8-bit algorithm is restated here in NumPy and pinned against
PIL itself (tests/test_oracle.py, and tests/golden/resize_golden.npz written by make_golden.py by calling PIL):
  * per output coordinate: centre = (xx + 0.5) * scale, support = 2 * max(scale, 1), window [xmin, xmin + n),
    bicubic (a = -0.5) weights evaluated in double, normalised to sum 1, converted to fixed point
    round-half-away(w * 2^22);
  * horizontal pass first (into a uint8 image), then vertical; each output = clip8((2^21 + sum pix * k) >> 22).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box: (bounds (out, 2) int, kk (out, ksize) int)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """One resampling pass along `axis` (0 = vertical, 1 = horizontal) of a (H, W, C) uint8 image."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((len(bounds),) + src.shape[1:], dtype=np.uint8)
    for xx, (xmin, n) in enumerate(bounds):
        acc = np.tensordot(kk[xx, :n].astype(np.int64), src[xmin:xmin + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bicubic(img, out_h, out_w):
    """PIL Image.resize((out_w, out_h), BICUBIC) of a (H, W, 3) uint8 image: horizontal pass, then vertical."""
    img = np.asarray(img, dtype=np.uint8)
    H, W = img.shape[:2]
    if out_w != W:
        img = _pass(img, *precompute_coeffs(W, out_w), axis=1)
    if out_h != H:
        img = _pass(img, *precompute_coeffs(H, out_h), axis=0)
    return img


def shortest_edge_size(H, W, size=224):
    """transformers get_resize_output_image_size(default_to_square=False): (new_h, new_w)."""
    short, long = (W, H) if W <= H else (H, W)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if W <= H else (new_short, new_long)


def clip_preprocess_u8(img, size=224):
    """resize shorter edge -> size (bicubic), centre crop size x size; the uint8 image the CLIP tower's rescale sees."""
    H, W = img.shape[:2]
    nh, nw = shortest_edge_size(H, W, size)
    r = resize_bicubic(img, nh, nw)
    top, left = (nh - size) // 2, (nw - size) // 2
    return r[top:top + size, left:left + size]
