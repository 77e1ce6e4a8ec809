"""Generates tests/golden/flava_pretraining_golden.pt by running the UNMODIFIED reference `FLAVAPretrainingLoss`
(torchmultimodal/modules/losses/flava.py:296-484) and `FLAVAForPreTraining` (models/flava/model.py:300-377, with the
deterministic stub codebook of tests/flava_pretraining_cases.py instead of DalleVAEEncoder) on seeded inputs.

    python tests/golden/make_flava_pretraining_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "iopath_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "tests"))

from torchmultimodal.models.flava.model import flava_model, FLAVAForPreTraining  # noqa: E402
from torchmultimodal.modules.losses.flava import FLAVAPretrainingLoss  # noqa: E402

import flava_pretraining_cases as PC  # noqa: E402


def main():
    torch.set_num_threads(8)
    out = {}
    loss = PC.build_loss(FLAVAPretrainingLoss)
    out["loss_param_checksum"] = PC.param_checksum(loss)
    for name, kw in PC.loss_calls().items():
        with torch.no_grad():
            o = loss(**kw)
        out[f"loss.{name}"] = PC.flatten_loss_output(o)
        print(name, {k: (tuple(v.shape) if v.dim() else round(float(v), 4)) for k, v in out[f"loss.{name}"].items()})
    m = PC.build_model(flava_model, FLAVAForPreTraining, FLAVAPretrainingLoss)
    inp, _ = PC.model_inputs()
    with torch.no_grad():
        o = m(**inp)
    out["model_param_checksum"] = PC.param_checksum(m)
    out["model"] = PC.flatten_loss_output(o)
    print("model", {k: (tuple(v.shape) if v.dim() else round(float(v), 4)) for k, v in out["model"].items()})
    path = os.path.join(HERE, "flava_pretraining_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
