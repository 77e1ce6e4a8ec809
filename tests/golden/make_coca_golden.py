"""Generates tests/golden/coca_golden.pt by running the UNMODIFIED reference CoCa (imported from /root/reference,
build container only) on the cases of tests/coca_cases.py.

    python tests/golden/make_coca_golden.py

Stored per case: inputs, parameter checksum (weights are re-created from seeds by the tests), the three tensors of
`CoCaModel.forward` and — for the parallel-pooler case, the only one the reference's CoCaForPretraining supports —
the contrastive and captioning losses.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "iopath_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "tests"))

from torchmultimodal.models.coca.coca_model import coca_for_pretraining  # noqa: E402

import coca_cases as CC  # noqa: E402


def main():
    torch.set_num_threads(8)
    out = {}
    for name, c in CC.CASES.items():
        m = CC.build(coca_for_pretraining, name)
        inp = CC.inputs(name)
        with torch.no_grad():
            mo = m.model(inp["images"], inp["texts"])
            rec = {"inputs": inp, "param_checksum": CC.param_checksum(m),
                   "image_pooled_output": mo.image_pooled_output.clone(), "text_pooled_output": mo.text_pooled_output.clone(),
                   "multimodal_embeddings": mo.multimodal_embeddings.clone()}
            if not c["kwargs"]["cascaded_pooler"]:
                losses = m(inp["images"], inp["texts"])
                rec["contrastive"], rec["captioning"] = losses["contrastive"].clone(), losses["captioning"].clone()
        out[name] = rec
        print(name, {k: (tuple(v.shape) if v.dim() else float(v)) for k, v in rec.items() if torch.is_tensor(v)})
    path = os.path.join(HERE, "coca_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
