"""Oracle (test infrastructure): 8x8x8 colour histograms, RGB and OpenCV-HSV.

PARITY UNPINNED BY THE REFERENCE: /root/reference contains no histogram code
(the nearest function is the k-means dominant colour at
imageProcessing.py:73-120, which is not on the path).  The definition used here
is the one BASELINE.json's north_star names (512-bin joint histogram of uint8
pixels = config.EMBEDDING_DIM) and it is pinned against OpenCV 4.13:

  RGB: bin = (r>>5)*64 + (g>>5)*8 + (b>>5)
       == cv2.calcHist([img],[0,1,2],None,[8,8,8],[0,256]*3)
  HSV: (h,s,v) = cv2.cvtColor(img, cv2.COLOR_RGB2HSV)  (8-bit, H in [0,180))
       bin = (h*8//180)*64 + (s>>5)*8 + (v>>5)
       == cv2.calcHist([hsv],[0,1,2],None,[8,8,8],[0,180,0,256,0,256])

`rgb_to_hsv_u8` is an integer restatement of OpenCV's 8-bit RGB2HSV fixed-point
code (published algorithm: modules/imgproc/src/color_hsv.simd.hpp, RGB2HSV_b,
hsv_shift = 12); tests/test_oracle.py checks it on all 2^24 colours.
"""
import numpy as np

BINS = 8
NBINS = BINS ** 3
HSV_SHIFT = 12


def _div_tables():
    i = np.arange(1, 256, dtype=np.float64)
    sdiv = np.zeros(256, dtype=np.int64)
    hdiv = np.zeros(256, dtype=np.int64)
    sdiv[1:] = np.rint((255 << HSV_SHIFT) / i).astype(np.int64)
    hdiv[1:] = np.rint((180 << HSV_SHIFT) / (6.0 * i)).astype(np.int64)
    return sdiv, hdiv


SDIV, HDIV = _div_tables()


def rgb_to_hsv_u8(rgb):
    """(…,3) uint8 RGB -> (…,3) uint8 HSV, bit-identical to cv2.cvtColor(COLOR_RGB2HSV)."""
    rgb = np.asarray(rgb, dtype=np.uint8)
    r = rgb[..., 0].astype(np.int64)
    g = rgb[..., 1].astype(np.int64)
    b = rgb[..., 2].astype(np.int64)
    v = np.maximum(np.maximum(r, g), b)
    mn = np.minimum(np.minimum(r, g), b)
    d = v - mn
    s = (d * SDIV[v] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT
    h0 = np.where(v == r, g - b, np.where(v == g, b - r + 2 * d, r - g + 4 * d))
    h = (h0 * HDIV[d] + (1 << (HSV_SHIFT - 1))) >> HSV_SHIFT      # arithmetic shift (floor)
    h = np.where(h < 0, h + 180, h)
    return np.stack([h, s, v], axis=-1).astype(np.uint8)


def bin_index(img, colorspace="rgb"):
    """Per-pixel bin id in [0, 512) for (…,3) uint8 pixels."""
    img = np.asarray(img, dtype=np.uint8)
    if colorspace == "rgb":
        c0 = img[..., 0].astype(np.int64) >> 5
    elif colorspace == "hsv":
        img = rgb_to_hsv_u8(img)
        c0 = (img[..., 0].astype(np.int64) * BINS) // 180
    else:
        raise ValueError(colorspace)
    c1 = img[..., 1].astype(np.int64) >> 5
    c2 = img[..., 2].astype(np.int64) >> 5
    return c0 * (BINS * BINS) + c1 * BINS + c2


def histogram(images, colorspace="rgb"):
    """(B,H,W,3) uint8 -> (B,512) uint32 counts."""
    images = np.asarray(images, dtype=np.uint8)
    if images.ndim == 3:
        images = images[None]
    B = images.shape[0]
    out = np.zeros((B, NBINS), dtype=np.uint32)
    for i in range(B):
        out[i] = np.bincount(bin_index(images[i], colorspace).ravel(), minlength=NBINS)
    return out


def histogram_cv2(image, colorspace="rgb"):
    """OpenCV reference for one (H,W,3) uint8 RGB image -> (512,) float32 counts."""
    import cv2
    image = np.ascontiguousarray(image, dtype=np.uint8)
    if colorspace == "rgb":
        h = cv2.calcHist([image], [0, 1, 2], None, [BINS] * 3, [0, 256, 0, 256, 0, 256])
    else:
        hsv = cv2.cvtColor(image, cv2.COLOR_RGB2HSV)
        h = cv2.calcHist([hsv], [0, 1, 2], None, [BINS] * 3, [0, 180, 0, 256, 0, 256])
    return h.reshape(-1)


def embedding(images, colorspace="rgb"):
    """Histogram embedding surface (shape of ImageEmbeddingSystem.generate_embedding,
    ImageEmbeddingSystem.py:85-94): raw vector = counts as fp32, magnitude = its L2 norm,
    returned as (unit vectors (B,512) fp32, magnitudes (B,) fp32)."""
    counts = histogram(images, colorspace).astype(np.float32)
    mag = np.array([np.linalg.norm(c) for c in counts], dtype=np.float32)
    return counts / mag[:, None], mag
