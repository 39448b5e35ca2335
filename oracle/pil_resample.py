"""CPU oracle — TEST INFRASTRUCTURE ONLY.  Restatement of Pillow's 8-bit-per-channel BICUBIC resampling
(src/libImaging/Resample.c of Pillow, the `ImagingResample` path that `Image.resize(..., resample=BICUBIC)` takes for
uint8 images), which is what the reference's CLIPProcessor runs on the host (lib/model_zoo/clip.py:88-94: tensor ->
ToPILImage -> CLIPProcessor resize (shortest side 224, bicubic) -> centre crop -> rescale -> normalise).

Third-party arithmetic: Pillow is not part of /root/reference (requirements.txt does not pin it; this image has 12.2.0).
Published algorithm restated here:
  * per output coordinate: support = 2 * max(scale, 1) taps around centre = (xx + 0.5) * scale, weights = Keys cubic
    (a = -0.5) of (x - centre + 0.5) / max(scale, 1), normalised to sum 1 in double precision;
  * weights -> int32 fixed point with 22 fractional bits, rounding half away from zero;
  * each pass accumulates int32 from 1 << 21, shifts right by 22 and clamps to [0, 255]; the HORIZONTAL pass runs first and
    its result is rounded to uint8 before the VERTICAL pass.
Pinned bit-exactly against Pillow itself in tests/test_preprocess_oracle.py (random sizes, up- and down-scaling)."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def coefficients(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc -> (bounds [out,2] int32 (xmin, count), kk [out, ksize] int32)."""
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


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resample_pass(img, bounds, kk, axis):
    """One 8-bit pass along `axis` (0 = vertical, 1 = horizontal) of an [H, W, C] uint8 image."""
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + img.shape[1:], dtype=np.uint8)
    for xx in range(bounds.shape[0]):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(n):
            acc += img[xmin + x] * int(kk[xx, x])
        out[xx] = _clip8(acc)
    return np.moveaxis(out, 0, axis)


def resize_bicubic_u8(img, out_w, out_h):
    """[H, W, C] uint8 -> [out_h, out_w, C] uint8, horizontal pass first (as ImagingResampleInner)."""
    h, w = img.shape[:2]
    if out_w != w:
        img = resample_pass(img, *coefficients(w, out_w), axis=1)
    if out_h != h:
        img = resample_pass(img, *coefficients(h, out_h), axis=0)
    return img


IMAGE_MEAN = (0.48145466, 0.4578275, 0.40821073)
IMAGE_STD = (0.26862954, 0.26130258, 0.27577711)


def clip_preprocess(images01, size=224):
    """[n, 3, H, W] float in [0,1] (numpy) -> fp32 [n, 3, size, size]: ToPILImage (x*255 truncated to uint8), resize the
    shortest side to `size` (long side int(size * long / short), transformers 4.24), centre crop, /255, normalise."""
    out = []
    for im in images01:
        u8 = (np.clip(im, 0.0, 1.0).astype(np.float32) * np.float32(255)).astype(np.uint8).transpose(1, 2, 0)
        h, w = u8.shape[:2]
        nw, nh = (size, int(size * h / w)) if w <= h else (int(size * w / h), size)
        r = resize_bicubic_u8(u8, nw, nh)
        left, top = (nw - size) // 2, (nh - size) // 2
        r = r[top:top + size, left:left + size].astype(np.float32) / np.float32(255.0)
        r = (r - np.asarray(IMAGE_MEAN, dtype=np.float32)) / np.asarray(IMAGE_STD, dtype=np.float32)
        out.append(r.transpose(2, 0, 1))
    return np.stack(out).astype(np.float32)
