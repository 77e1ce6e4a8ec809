"""CPU: the optimizer oracle (oracle/optim_oracle.py) against tensors produced by the UNMODIFIED reference
AnyPrecisionAdamW (tests/golden/anyprecision_golden.pt), and the fused kernel's arithmetic — restated op for op in
torch fp32 — against the same goldens (catches a wrong rounding order before any GPU time is spent)."""
import os

import torch

from oracle import optim_oracle as OO

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "anyprecision_golden.pt"), weights_only=False)


def test_optimizer_oracle_matches_reference_bit_for_bit():
    for name, c in G["cases"].items():
        cfg = c["config"]
        p = G["p0"].clone()
        st = OO.new_state(p, cfg["momentum_dtype"], cfg["variance_dtype"], cfg["compensation_dtype"])
        for g in G["grads"]:
            OO.anyprecision_adamw_step(p, g, st, lr=cfg["lr"], betas=cfg["betas"], eps=cfg["eps"],
                                       weight_decay=cfg["weight_decay"], use_kahan_summation=cfg["use_kahan_summation"])
        assert torch.equal(p, c["p"]), name
        assert torch.equal(st["exp_avg"], c["exp_avg"]) and torch.equal(st["exp_avg_sq"], c["exp_avg_sq"]), name
        if cfg["use_kahan_summation"]:
            assert torch.equal(st["compensation"], c["compensation"]), name


def _rn(x, dt):
    return x.to(dt).float()


def kernel_formula(p, g, m, v, comp, cfg, step):
    """anyprecision_adamw_kernel (csrc/elementwise.cu), one element-wise pass in fp32 with the kernel's op order."""
    md, vd, cd = cfg["momentum_dtype"], cfg["variance_dtype"], cfg["compensation_dtype"]
    lr, (b1, b2), eps, wd = cfg["lr"], cfg["betas"], cfg["eps"], cfg["weight_decay"]
    f = lambda x: torch.tensor(x, dtype=torch.float32)   # noqa: E731
    b1f, b2f = f(b1), f(b2)
    bc1 = 1 - torch.pow(b1f, f(float(step)))
    nss = -((1 / bc1) * f(lr))        # Tensor.__rtruediv__: reciprocal * scalar
    dc = torch.sqrt(1 - torch.pow(b2f, f(float(step))))
    decay = f(1.0 - lr * wd) if wd else f(1.0)
    a1, a2 = f(1.0 - b1), f(1.0 - b2)
    P = p * decay
    mb = _rn(m.float() * b1f, md)
    M = _rn((a1.double() * g.double() + mb.double()).float(), md)          # fmaf
    V = _rn(((a2 * g).double() * g.double() + _rn(v.float() * b2f, vd).double()).float(), vd)   # fmaf
    cv = _rn(_rn(_rn(torch.sqrt(V), vd) / dc, vd) + f(eps), vd)
    upd = (nss * M) / cv
    if comp is not None:
        C = _rn(comp.float() + upd, cd)
        T = P
        P = P + C
        C = _rn(C + (T - P), cd)
        comp = C.to(cd)
    else:
        P = P + upd
    return P, M.to(md), V.to(vd), comp


def test_kernel_arithmetic_restated_matches_reference():
    for name, c in G["cases"].items():
        cfg = c["config"]
        p = G["p0"].clone()
        m = torch.zeros_like(p, dtype=cfg["momentum_dtype"])
        v = torch.zeros_like(p, dtype=cfg["variance_dtype"])
        comp = torch.zeros_like(p, dtype=cfg["compensation_dtype"]) if cfg["use_kahan_summation"] else None
        for t, g in enumerate(G["grads"], 1):
            p, m, v, comp = kernel_formula(p, g, m, v, comp, cfg, t)
        # with the reference's contractions reproduced (add_(alpha) and addcmul_ are FMAs, lr / tensor is
        # reciprocal * lr, addcdiv_ is (value * t1) / t2) every tensor matches the reference BIT FOR BIT
        assert torch.equal(m.float(), c["exp_avg"].float()) and torch.equal(v.float(), c["exp_avg_sq"].float()), name
        assert torch.equal(p, c["p"]), name
        if comp is not None:
            assert torch.equal(comp.float(), c["compensation"].float()), name
