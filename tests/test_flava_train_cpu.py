"""Host logic of the FLAVA training runtime (engine_flava_train.py + engine.TransformerStack) WITHOUT a GPU: the
kernel wrappers are swapped for their torch emulation (tests/emu_ops.py), so what is checked here is the schedule —
buffer routing, packed q/k/v parameter views, per-call saved activations (the same encoder runs twice before the
backward), gradient slots — against autograd over the fp32 oracle.  The kernels themselves are checked on the GPU
(tests/test_gpu_flava_train.py)."""
import math

import pytest
import torch

import emu_ops
import flava_cases as FC
import test_gpu_flava_train as G   # shared helpers only (its tests carry the gpu marker)


@pytest.fixture()
def emu(monkeypatch):
    emu_ops.install(monkeypatch)


def test_flava_training_schedule_against_oracle_with_emulated_kernels(emu):
    from multimodal_b200.models.flava import flava_model

    name = "flava_small"
    m = FC.build(flava_model, name)
    cpu = torch.device("cpu")
    m = G._grad_parity(cpu, m, G._cfg(FC.CASES[name]["kwargs"]), FC.inputs(name), "cpu_emu", 6e-2)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.01 * torch.randn_like(p))
            p.grad = None
    G._grad_parity(cpu, m, G._cfg(FC.CASES[name]["kwargs"]), FC.inputs(name), "cpu_emu_step2", 6e-2)


def test_flava_mm_direct_call_and_frozen_encoder_with_emulated_kernels(emu):
    G.test_flava_mm_encoder_direct_call_and_frozen_parts.__wrapped__(torch.device("cpu")) \
        if hasattr(G.test_flava_mm_encoder_direct_call_and_frozen_parts, "__wrapped__") \
        else G.test_flava_mm_encoder_direct_call_and_frozen_parts(torch.device("cpu"))
