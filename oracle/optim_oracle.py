"""TEST INFRASTRUCTURE — CPU restatement of the reference optimizer on the path's caller side.

`anyprecision_adamw_step` restates torchmultimodal/modules/optimizers/anyprecision.py:139-199 for one parameter
tensor with plain torch ops (same in-place op sequence, hence the same roundings to the state dtypes).  Pinned against
tensors produced by the unmodified reference (tests/golden/anyprecision_golden.pt, made by
tests/golden/make_anyprecision_golden.py); only tests/ may import it.
"""
import torch


def anyprecision_adamw_step(p, grad, state, *, lr, betas, eps, weight_decay, use_kahan_summation):
    """In-place update of `p` and of state = {"step", "exp_avg", "exp_avg_sq"[, "compensation"]} (anyprecision.py:155-199)."""
    beta1, beta2 = betas
    state["step"] += 1                                     # :155
    step = state["step"]
    exp_avg, exp_avg_sq = state["exp_avg"], state["exp_avg_sq"]
    if weight_decay:                                       # :164-165
        p.mul_(1 - lr * weight_decay)
    exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)        # :168
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)   # :171
    bias_correction1 = 1 - beta1 ** step                   # :174
    step_size = lr / bias_correction1                      # :176
    denom_correction = (1 - beta2 ** step) ** 0.5          # :179
    centered_variance = (exp_avg_sq.sqrt() / denom_correction).add_(eps)   # :181
    if use_kahan_summation:                                # :184-194
        comp = state["compensation"]
        comp.addcdiv_(exp_avg, centered_variance, value=-step_size)
        temp = p.detach().clone()
        p.add_(comp)
        comp.add_(temp.sub_(p))
    else:
        p.addcdiv_(exp_avg, centered_variance, value=-step_size)   # :198


def new_state(p, momentum_dtype, variance_dtype, compensation_dtype=None):
    st = {"step": torch.tensor(0.0), "exp_avg": torch.zeros_like(p, dtype=momentum_dtype),
          "exp_avg_sq": torch.zeros_like(p, dtype=variance_dtype)}
    if compensation_dtype is not None:
        st["compensation"] = torch.zeros_like(p, dtype=compensation_dtype)
    return st
