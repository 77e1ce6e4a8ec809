"""`CLIPImageTransform` — drop-in for the image half of torchmultimodal/transforms/clip_transform.py:300-352, on the GPU.

Same constructor (`image_size`, `image_interpolation`, `image_mean`, `image_std`, `is_train`) and the same result, bit for
bit, as the reference's PIL / torchvision pipeline — Resize(BICUBIC) + CenterCrop (eval) or RandomResizedCrop (train),
RGB, ToTensor, Normalize — but computed by one fused resample kernel on DECODED images (`mmb_clip_image_transform`):
    inputs : a PIL image / HWC uint8 tensor (CPU or CUDA), or a list of them (sizes may differ);
    output : fp32 CUDA tensor [3, S, S] (single image) or [N, 3, S, S], ready for `CLIPViTEncoder`.
In train mode the crop boxes are sampled on the host exactly as `torchvision.transforms.RandomResizedCrop.get_params` does
(same torch RNG calls, so a seeded run picks the reference's crops); only the pixels are processed on the device.
JPEG decoding is not part of the transform (as in the reference, which receives decoded PIL images).
The text half (`CLIPTextTransform`, `CLIPBPETransform`: ftfy + regex + a BPE merges file that the reference downloads) is
host string processing with no device work to move; it is not provided here.
"""
import math
from typing import List, Sequence, Tuple, Union

import torch
from torch import nn, Tensor

from .. import ops
from .._lib import MMBError

CLIP_DEFAULT_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_DEFAULT_STD = (0.26862954, 0.26130258, 0.27577711)


def _is_bicubic(mode) -> bool:
    name = getattr(mode, "value", mode)
    return str(name).lower() in ("bicubic", "interpolationmode.bicubic", "3")


def _to_u8_hwc(image, device) -> Tensor:
    """PIL image or uint8 tensor [H, W, 3] -> contiguous CUDA uint8 [H, W, 3]."""
    if isinstance(image, Tensor):
        t = image
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[-1] != 3:
            raise MMBError("CLIPImageTransform: tensors must be uint8 [H, W, 3] (decoded RGB, channels last)")
    else:   # PIL.Image (duck-typed: no hard dependency on Pillow in the product)
        import numpy as np
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(image.convert("RGB"))))
    return t.to(device, non_blocking=True).contiguous()


def random_resized_crop_params(height: int, width: int, scale=(0.08, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)) -> Tuple[int, int, int, int]:
    """torchvision.transforms.RandomResizedCrop.get_params (transforms.py:936-977), same torch RNG consumption."""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,)).item()
            j = torch.randint(0, width - w + 1, size=(1,)).item()
            return i, j, h, w
    in_ratio = float(width) / float(height)   # fallback: central crop
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


class CLIPImageTransform(nn.Module):
    def __init__(self, image_size: Union[int, Tuple[int, int]] = 224, image_interpolation="bicubic",
                 image_mean: Tuple[float, float, float] = CLIP_DEFAULT_MEAN,
                 image_std: Tuple[float, float, float] = CLIP_DEFAULT_STD, is_train: bool = True,
                 device: Union[str, torch.device] = "cuda") -> None:
        super().__init__()
        if isinstance(image_size, (tuple, list)):
            if len(image_size) != 2 or image_size[0] != image_size[1]:
                raise NotImplementedError("only square output sizes are on the accelerated path")
            image_size = image_size[0]
        if not _is_bicubic(image_interpolation):
            raise NotImplementedError("only bicubic interpolation (the reference default) is on the accelerated path")
        self.image_size = int(image_size)
        self.image_mean, self.image_std = tuple(image_mean), tuple(image_std)
        self.is_train = bool(is_train)
        self.device = torch.device(device)

    # -- geometry (host) ------------------------------------------------------------------------------------------
    def _geometry(self, H: int, W: int, pitch: int) -> List[int]:
        S = self.image_size
        if self.is_train:
            top, left, h, w = random_resized_crop_params(H, W)
            box, rw, rh, cl, ct = (left, top, w, h), S, S, 0, 0
        else:   # torchvision Resize(int): short edge -> S, long edge -> int(S * long / short); then CenterCrop(S)
            short, long = (W, H) if W <= H else (H, W)
            new_long = int(S * long / short)
            rw, rh = (S, new_long) if W <= H else (new_long, S)
            if rw < S or rh < S:
                raise NotImplementedError("CenterCrop padding (image smaller than the crop) is not on the accelerated path")
            box = (0, 0, W, H)
            ct, cl = int(round((rh - S) / 2.0)), int(round((rw - S) / 2.0))
        for n_in, n_out in ((box[2], rw), (box[3], rh)):
            taps = int(math.ceil(2.0 * max(n_in / n_out, 1.0))) * 2 + 1
            if taps > ops.clip_image_transform_max_taps():
                raise NotImplementedError(f"down-scaling {n_in} -> {n_out} needs {taps} filter taps per pixel (limit "
                                          f"{ops.clip_image_transform_max_taps()}); pre-scale such images on the host")
        flags = (1 if rw != box[2] else 0) | (2 if rh != box[3] else 0)
        return [H, W, pitch, box[0], box[1], box[2], box[3], rw, rh, cl, ct, flags]

    def forward(self, image) -> Tensor:
        single = not isinstance(image, (list, tuple))
        images: Sequence = [image] if single else list(image)
        if not images:
            raise ValueError("CLIPImageTransform: empty batch")
        dev = self.device
        srcs = [_to_u8_hwc(im, dev) for im in images]
        geom = [self._geometry(t.shape[0], t.shape[1], t.stride(0)) for t in srcs]
        S = self.image_size
        ptrs = torch.tensor([t.data_ptr() for t in srcs], dtype=torch.int64).to(dev)
        g = torch.tensor(geom, dtype=torch.int32).to(dev)
        out = torch.empty((len(srcs), 3, S, S), device=dev, dtype=torch.float32)
        ops.clip_image_transform(ptrs, g, out, self.image_mean, self.image_std)
        # `srcs` may be freed when this frame returns: the caching allocator only reuses their memory for work queued
        # later on the same stream, i.e. after the kernel has read them
        return out[0] if single else out
