"""GPU parity of the CLIP image transform (image half): BIT-EXACT against the oracle (which is pinned
bit-exact to the reference's PIL / torchvision pipeline) and against the reference golden vectors."""
import os

import numpy as np
import pytest
import torch

from oracle import clip_transform_oracle as CT

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "clip_transform_golden.pt")


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def test_clip_image_transform_reference_golden_bit_exact(dev):
    from multimodal_b200.transforms.clip_transform import CLIPImageTransform

    g = torch.load(GOLD)
    S = g["size"]
    ev = CLIPImageTransform(image_size=S, is_train=False)
    out = ev([im for im in g["images"]])
    assert out.shape == g["eval"].shape and out.dtype == torch.float32 and out.is_cuda
    assert torch.equal(out.cpu(), g["eval"])
    assert torch.equal(ev(g["images"][1]).cpu(), g["eval"][1])           # single image -> [3, S, S]
    tr = CLIPImageTransform(image_size=S, is_train=True)
    torch.manual_seed(g["train_seed"])
    out = torch.stack([tr(im) for im in g["images"]])                      # one sampler call per image, as the reference
    assert torch.equal(out.cpu(), g["train"])


@pytest.mark.parametrize("H,W,S", [(300, 400, 224), (231, 229, 224), (100, 150, 64), (500, 333, 96), (224, 300, 224),
                                   (64, 1000, 48), (1080, 1920, 224), (224, 224, 224)])
def test_clip_image_transform_eval_bit_exact_against_oracle(dev, H, W, S):
    from multimodal_b200.transforms.clip_transform import CLIPImageTransform

    rng = np.random.default_rng(H * 7 + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    out = CLIPImageTransform(image_size=S, is_train=False)(torch.from_numpy(img).to(dev))
    assert torch.equal(out.cpu(), torch.from_numpy(CT.clip_image_transform_eval(img, S)))


def test_clip_image_transform_train_batch_bit_exact_against_oracle(dev):
    from multimodal_b200.transforms.clip_transform import CLIPImageTransform, random_resized_crop_params

    rng = np.random.default_rng(3)
    sizes = [(333, 500), (480, 640), (200, 200), (97, 301), (512, 384), (225, 226)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    t = CLIPImageTransform(image_size=224, is_train=True)
    torch.manual_seed(11)
    out = t([torch.from_numpy(i) for i in imgs])                            # CPU tensors: copied to the device inside
    torch.manual_seed(11)
    for i, im in enumerate(imgs):
        crop = random_resized_crop_params(im.shape[0], im.shape[1])
        assert torch.equal(out[i].cpu(), torch.from_numpy(CT.clip_image_transform_crop(im, crop, 224))), i
    # properties at the benchmark's batch shape: finite, per-channel range of a normalised uint8 image
    big = t([torch.from_numpy(imgs[0])] * 64)
    assert big.shape == (64, 3, 224, 224) and torch.isfinite(big).all()
    lo = torch.tensor([(0 - m) / s for m, s in zip(CT.CLIP_DEFAULT_MEAN, CT.CLIP_DEFAULT_STD)], device=dev)
    hi = torch.tensor([(1 - m) / s for m, s in zip(CT.CLIP_DEFAULT_MEAN, CT.CLIP_DEFAULT_STD)], device=dev)
    assert (big.amin((0, 2, 3)) >= lo - 1e-6).all() and (big.amax((0, 2, 3)) <= hi + 1e-6).all()
