"""Host logic of the CoCa training runtime (engine_coca_train.py + the cross-attention / general-mask extensions of
engine.TransformerStack) WITHOUT a GPU: kernel wrappers swapped for their torch emulation (tests/emu_ops.py), the result
compared with autograd over the fp32 oracle (oracle/coca_oracle.py).  Kernels proper: tests/test_gpu_coca_train.py."""
import pytest
import torch

import coca_cases as CC
import emu_ops
import test_gpu_coca_train as G


@pytest.fixture()
def emu(monkeypatch):
    emu_ops.install(monkeypatch)


@pytest.mark.parametrize("name", list(CC.CASES))
def test_coca_training_schedule_against_oracle_with_emulated_kernels(emu, name):
    G.coca_grad_parity(torch.device("cpu"), name, "cpu_emu_" + name, with_contrastive=False)


@pytest.mark.parametrize("masked", [False, True])
def test_standalone_encoder_layers_train_with_emulated_kernels(emu, masked):
    G.standalone_layers_grad_parity(torch.device("cpu"), masked)
