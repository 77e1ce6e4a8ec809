"""Generates tests/golden/anyprecision_golden.pt by running the UNMODIFIED reference optimizer
(torchmultimodal/modules/optimizers/anyprecision.py) in the build container (CPU).  Not used at test time.

    PYTHONPATH=tests/golden/_shim:/root/reference python tests/golden/make_anyprecision_golden.py
"""
import importlib.util
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location(
    "ref_anyprecision", "/root/reference/torchmultimodal/modules/optimizers/anyprecision.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)   # the file imports torch only: no package import (and no iopath) needed

CASES = {
    # name: (momentum, variance, kahan, compensation, lr, betas, eps, wd)
    "fp32_states": (torch.float32, torch.float32, False, None, 1e-3, (0.9, 0.999), 1e-8, 0.01),
    "default_bf16_variance": (torch.float32, torch.bfloat16, False, None, 1e-3, (0.9, 0.999), 1e-8, 0.0),
    "bf16_both_kahan_bf16": (torch.bfloat16, torch.bfloat16, True, torch.bfloat16, 5e-4, (0.9, 0.98), 1e-6, 0.2),
    "fp32_m_kahan_fp32": (torch.float32, torch.bfloat16, True, torch.float32, 1e-3, (0.9, 0.999), 1e-8, 0.1),
}
torch.manual_seed(7)
N, STEPS = 4096, 3
p0 = torch.randn(N) * 0.05
grads = [torch.randn(N) * (0.02 if i else 0.5) for i in range(STEPS)]
out = {"p0": p0, "grads": grads, "cases": {}}
for name, (md, vd, kahan, cd, lr, betas, eps, wd) in CASES.items():
    p = torch.nn.Parameter(p0.clone())
    kw = dict(lr=lr, betas=betas, eps=eps, weight_decay=wd, use_kahan_summation=kahan, momentum_dtype=md, variance_dtype=vd)
    if cd is not None:
        kw["compensation_buffer_dtype"] = cd
    opt = mod.AnyPrecisionAdamW([p], **kw)
    for g in grads:
        p.grad = g.clone()
        opt.step()
    st = opt.state[p]
    out["cases"][name] = {"config": dict(momentum_dtype=md, variance_dtype=vd, use_kahan_summation=kahan,
                                         compensation_dtype=cd, lr=lr, betas=betas, eps=eps, weight_decay=wd),
                          "p": p.detach().clone(), "exp_avg": st["exp_avg"].clone(), "exp_avg_sq": st["exp_avg_sq"].clone(),
                          "compensation": st["compensation"].clone() if kahan else None}
torch.save(out, os.path.join(HERE, "anyprecision_golden.pt"))
print("wrote", {k: float(v["p"].abs().sum()) for k, v in out["cases"].items()})
