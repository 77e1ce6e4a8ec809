"""Pins oracle/clip_transform_oracle.py (the numpy restatement of Pillow's resample + torchvision's size arithmetic) —
bit-exact against the committed golden vectors produced by the reference's own transform stack, and against PIL /
torchvision run here on random images — and checks the host half of the drop-in (`random_resized_crop_params` consumes
the torch RNG exactly like torchvision's RandomResizedCrop.get_params; geometry rows)."""
import os

import numpy as np
import pytest
import torch

from oracle import clip_transform_oracle as CT

GOLD = os.path.join(os.path.dirname(__file__), "golden", "clip_transform_golden.pt")


def test_oracle_reproduces_reference_golden_bit_exact():
    from multimodal_b200.transforms.clip_transform import random_resized_crop_params

    g = torch.load(GOLD)
    S = g["size"]
    for img, ref in zip(g["images"], g["eval"]):
        assert np.array_equal(CT.clip_image_transform_eval(img.numpy(), S), ref.numpy())
    torch.manual_seed(g["train_seed"])
    for img, ref in zip(g["images"], g["train"]):
        crop = random_resized_crop_params(img.shape[0], img.shape[1])
        assert np.array_equal(CT.clip_image_transform_crop(img.numpy(), crop, S), ref.numpy())


def test_oracle_against_pillow_and_torchvision_on_random_sizes():
    Image = pytest.importorskip("PIL.Image")
    T = pytest.importorskip("torchvision.transforms")
    from torchvision.transforms import InterpolationMode

    rng = np.random.default_rng(1)
    for (H, W, S) in [(300, 400, 224), (231, 229, 224), (100, 150, 64), (500, 333, 96), (224, 300, 224), (64, 1000, 48)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        ref = T.Compose([T.Resize(S, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(S), T.ToTensor(),
                         T.Normalize(CT.CLIP_DEFAULT_MEAN, CT.CLIP_DEFAULT_STD)])(Image.fromarray(img)).numpy()
        assert np.array_equal(CT.clip_image_transform_eval(img, S), ref), (H, W, S)


def test_crop_sampler_matches_torchvision_rng_stream():
    T = pytest.importorskip("torchvision.transforms")
    from multimodal_b200.transforms.clip_transform import random_resized_crop_params

    img = torch.zeros(3, 123, 77)
    torch.manual_seed(5)
    want = [T.RandomResizedCrop.get_params(img, (0.08, 1.0), (3.0 / 4.0, 4.0 / 3.0)) for _ in range(20)]
    torch.manual_seed(5)
    got = [random_resized_crop_params(123, 77) for _ in range(20)]
    assert got == [tuple(w) for w in want]



def test_geometry_rows_and_limits():
    from multimodal_b200 import ops
    from multimodal_b200.transforms.clip_transform import CLIPImageTransform

    t = CLIPImageTransform(image_size=224, is_train=False, device="cpu")
    assert t._geometry(300, 400, 1200) == [300, 400, 1200, 0, 0, 400, 300, 298, 224, 37, 0, 3]
    assert t._geometry(224, 300, 900)[7:] == [300, 224, 38, 0, 0]          # nothing to resample: both passes skipped
    assert t._geometry(448, 224, 672)[7:] == [224, 448, 0, 112, 0]
    assert t._geometry(100, 100, 300)[5:] == [100, 100, 224, 224, 0, 0, 3]   # up-scaling: short edge -> 224
    with pytest.raises(NotImplementedError):
        CLIPImageTransform(image_interpolation="bilinear")
    assert ops.clip_image_transform_max_taps() == 64
    with pytest.raises(NotImplementedError):
        t._geometry(224 * 17, 224 * 17, 3 * 224 * 17)                         # 17x down-scaling: more than 64 taps
