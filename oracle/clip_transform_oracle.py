"""CPU restatement (numpy, integer / IEEE-double arithmetic) of the CLIP image transform — TEST INFRASTRUCTURE ONLY.

What the reference runs (torchmultimodal/transforms/clip_transform.py:300-352): on a PIL image,
    eval : torchvision Resize(size, BICUBIC) -> CenterCrop(size) -> convert("RGB") -> ToTensor() -> Normalize(mean, std)
    train: torchvision RandomResizedCrop(size, BICUBIC)          -> convert("RGB") -> ToTensor() -> Normalize(mean, std)
The resampling itself lives in a third-party dependency that is not vendored in /root/reference: **Pillow** (pinned by
this image: 12.2.0; the algorithm has not changed since 7.x), `src/libImaging/Resample.c`:
    precompute_coeffs     -> `resample_coeffs`   (double arithmetic, window rounding by C truncation)
    normalize_coeffs_8bpc -> 22-bit fixed point   (PRECISION_BITS = 32 - 8 - 2)
    ImagingResampleHorizontal_8bpc / Vertical_8bpc -> two passes with a uint8-rounded intermediate image
and torchvision (0.26) supplies the size arithmetic: Resize(int) -> short edge = size, long edge =
int(size * long / short) (transforms/functional.py:_compute_resized_output_size); CenterCrop offsets
int(round((h - ch) / 2.0)) (Python banker's rounding); ToTensor = uint8 / 255 in fp32; Normalize = (x - mean) / std.
Pinned by tests/test_clip_transform_cpu.py: bit-exact against PIL + torchvision run here, on random images and sizes, and
against the committed golden vectors (tests/golden/clip_transform_golden.pt).
"""
import math
from typing import Optional, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2
CLIP_DEFAULT_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_DEFAULT_STD = (0.26862954, 0.26130258, 0.27577711)


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def resample_coeffs(in_size: int, in0: float, in1: float, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc -> (bounds int32 [out, 2], kk int32 [out, ksize])."""
    scale = float(np.float32(in1) - np.float32(in0)) / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)         # C (int) truncates toward zero, as Python int() does
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bicubic_u8(img: np.ndarray, out_w: int, out_h: int, box: Optional[Tuple[int, int, int, int]] = None) -> np.ndarray:
    """PIL `Image.resize((out_w, out_h), BICUBIC, box=(left, top, right, bottom))` for an HWC uint8 image."""
    H, W, _ = img.shape
    left, top, right, bottom = box if box is not None else (0, 0, W, H)
    need_h = out_w != W or left != 0 or right != out_w
    need_v = out_h != H or top != 0 or bottom != out_h
    if not need_h and not need_v:
        return img.copy()
    bh, kh = resample_coeffs(W, left, right, out_w)
    bv, kv = resample_coeffs(H, top, bottom, out_h)
    src = img.astype(np.int64)
    if need_h:
        first, last = int(bv[0, 0]), int(bv[out_h - 1, 0] + bv[out_h - 1, 1])
        tmp = np.zeros((last - first, out_w, img.shape[2]), dtype=np.uint8)
        for x in range(out_w):
            x0, n = int(bh[x, 0]), int(bh[x, 1])
            acc = (src[first:last, x0:x0 + n, :] * kh[x, :n].astype(np.int64)[None, :, None]).sum(1) + (1 << (PRECISION_BITS - 1))
            tmp[:, x, :] = _clip8(acc)
        src = tmp.astype(np.int64)
        bv = bv.copy()
        bv[:, 0] -= first
    if not need_v:
        return src.astype(np.uint8)
    out = np.zeros((out_h, src.shape[1], img.shape[2]), dtype=np.uint8)
    for y in range(out_h):
        y0, n = int(bv[y, 0]), int(bv[y, 1])
        acc = (src[y0:y0 + n] * kv[y, :n].astype(np.int64)[:, None, None]).sum(0) + (1 << (PRECISION_BITS - 1))
        out[y] = _clip8(acc)
    return out


def resized_size(h: int, w: int, size: int) -> Tuple[int, int]:
    """torchvision Resize(int): (new_h, new_w) with the short edge = size."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def center_crop_offsets(h: int, w: int, ch: int, cw: int) -> Tuple[int, int]:
    return int(round((h - ch) / 2.0)), int(round((w - cw) / 2.0))


def to_tensor_normalize(img_u8: np.ndarray, mean=CLIP_DEFAULT_MEAN, std=CLIP_DEFAULT_STD) -> np.ndarray:
    x = img_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)          # ToTensor: fp32 division
    m = np.asarray(mean, dtype=np.float32)[:, None, None]
    s = np.asarray(std, dtype=np.float32)[:, None, None]
    return (x - m) / s                                                             # Normalize: sub then div, fp32


def clip_image_transform_eval(img: np.ndarray, size: int = 224, mean=CLIP_DEFAULT_MEAN, std=CLIP_DEFAULT_STD) -> np.ndarray:
    """CLIPImageTransform(is_train=False) on an HWC uint8 RGB image whose edges are >= size after the resize."""
    H, W, _ = img.shape
    nh, nw = resized_size(H, W, size)
    r = resize_bicubic_u8(img, nw, nh)
    top, left = center_crop_offsets(nh, nw, size, size)
    return to_tensor_normalize(r[top:top + size, left:left + size], mean, std)


def clip_image_transform_crop(img: np.ndarray, crop: Tuple[int, int, int, int], size: int = 224, mean=CLIP_DEFAULT_MEAN,
                              std=CLIP_DEFAULT_STD) -> np.ndarray:
    """CLIPImageTransform(is_train=True) given RandomResizedCrop's sampled (top, left, height, width): the crop is
    resized as a standalone image (torchvision F.resized_crop = crop, then resize)."""
    t, l, h, w = crop
    return to_tensor_normalize(resize_bicubic_u8(np.ascontiguousarray(img[t:t + h, l:l + w]), size, size), mean, std)
