"""Generates tests/golden/clip_transform_golden.pt: outputs of the REFERENCE image pipeline — the exact transform stack of
torchmultimodal/transforms/clip_transform.py:326-346 (torchvision Resize(BICUBIC) + CenterCrop / RandomResizedCrop, convert
RGB, ToTensor, Normalize on PIL images).  The reference module itself cannot be imported here (it imports `ftfy`, absent
from this image, at module level for its text half), so the stack is built from the same torchvision / Pillow calls the
reference makes.  Run: python tests/golden/make_clip_transform_golden.py"""
import os

import numpy as np
import torch
from PIL import Image
from torchvision import transforms as T
from torchvision.transforms import InterpolationMode

MEAN = (0.48145466, 0.4578275, 0.40821073)
STD = (0.26862954, 0.26130258, 0.27577711)


def main():
    rng = np.random.default_rng(7)
    S = 32
    joint = [lambda im: im.convert("RGB"), T.ToTensor(), T.Normalize(MEAN, STD)]
    evalt = T.Compose([T.Resize(S, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(S)] + joint)
    traint = T.Compose([T.RandomResizedCrop(S, interpolation=InterpolationMode.BICUBIC)] + joint)
    images, ev, tr = [], [], []
    for (H, W) in [(60, 80), (97, 51), (40, 40), (33, 200), (32, 47)]:
        base = rng.integers(0, 256, (H // 4 + 2, W // 4 + 2, 3), dtype=np.uint8)
        img = np.asarray(Image.fromarray(base).resize((W, H), Image.BILINEAR))      # some low-frequency content
        img = np.clip(img.astype(np.int32) + rng.integers(-20, 21, img.shape), 0, 255).astype(np.uint8)
        images.append(torch.from_numpy(img.copy()))
        pil = Image.fromarray(img)
        ev.append(evalt(pil))
    torch.manual_seed(123)
    for t in images:
        tr.append(traint(Image.fromarray(t.numpy())))
    out = dict(size=S, images=images, eval=torch.stack(ev), train=torch.stack(tr), train_seed=123)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "clip_transform_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
