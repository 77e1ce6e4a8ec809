"""CPU, world_size 2 over gloo: (1) the mirrored utils.distributed.gather_tensor keeps the reference's three autograd
modes (tests/utils/test_distributed.py:24-95 needs GPUs upstream and is therefore never run in the reference CI);
(2) the 'no gradient traffic' identity behind the CUDA distributed loss — gradients rebuilt from the local logits
block + the peers' row-LSE vectors equal autograd through the reference's all_gather-with-backprop formulation, for
all three BackpropTypes; (3) host-side layout / chunking logic of the symmetric buffers and the trainer."""
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _worker_gather(rank, world, port, q):
    from multimodal_b200.utils.distributed import BackpropType, concat_gather_all_gpu, gather_tensor, get_rank

    _init(rank, world, port)
    try:
        assert get_rank() == rank
        x = (torch.ones(3, 2) * (rank + 1)).requires_grad_(True)
        for mode in BackpropType:
            out = gather_tensor(x * 1.0, mode)
            assert len(out) == world
            for i, t in enumerate(out):
                assert torch.equal(t.detach(), torch.ones(3, 2) * (i + 1))
                has = t.grad_fn is not None
                assert has == (mode == BackpropType.GLOBAL or (mode == BackpropType.LOCAL and i == rank)), (mode, i)
        assert concat_gather_all_gpu(x, BackpropType.NONE).shape == (3 * world, 2)
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def _worker_lse(rank, world, port, q):
    from oracle import clip_oracle as O

    _init(rank, world, port)
    try:
        torch.manual_seed(100 + rank)
        B, E = 6, 16
        a0 = O.normalize(torch.randn(B, E, dtype=torch.float64))
        b0 = O.normalize(torch.randn(B, E, dtype=torch.float64))
        masks = (None, torch.tensor([1, 0, 1, 1, 0, 1], dtype=torch.bool) if rank == 0
                 else torch.tensor([0, 1, 1, 0, 0, 0], dtype=torch.bool))   # different row counts per rank
        for mode in ("GLOBAL", "LOCAL", "NONE"):
            for eps, mask in ((0.0, None), (0.1, None), (0.0, masks[1]), (0.1, masks[1])):
                a = a0.clone().requires_grad_(True)
                b = b0.clone().requires_grad_(True)
                s = torch.tensor(math.log(1 / 0.07), dtype=torch.float64, requires_grad=True)
                loss = O.contrastive_loss_distributed(a, b, s, mode, eps, mask)[0]
                loss.backward()
                l2, dA, dB, dS = O.contrastive_grads_lse_exchange(a0, b0, s.detach(), mode, eps, mask)
                torch.testing.assert_close(l2, loss.detach(), rtol=1e-12, atol=1e-12)
                torch.testing.assert_close(dA, a.grad, rtol=1e-10, atol=1e-12)
                torch.testing.assert_close(dB, b.grad, rtol=1e-10, atol=1e-12)
                torch.testing.assert_close(dS, s.grad, rtol=1e-10, atol=1e-12)
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=fn, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, "ok") for r in range(world)]
    assert all(p.exitcode == 0 for p in procs)


def test_gather_tensor_backprop_modes_gloo():
    _run(_worker_gather)


def test_lse_exchange_equals_allgather_backprop_gloo():
    _run(_worker_lse)


def test_symmetric_slot_layout():
    from multimodal_b200.symm import _Slots

    B, E = 1024, 512
    n = _Slots.size(B, E)
    raw = torch.zeros(n, dtype=torch.uint8)
    s = _Slots(raw, B, E)
    assert s.nbytes == n
    ptrs = [t.data_ptr() - raw.data_ptr() for p in range(2) for t in (s.a[p], s.b[p], s.lse_a[p], s.lse_b[p])]
    assert all(x % 256 == 0 for x in ptrs) and ptrs == sorted(ptrs) and len(set(ptrs)) == 8
    assert s.a[1].shape == (B, E) and s.a[1].dtype == torch.bfloat16 and s.lse_b[0].dtype == torch.float32
    assert s.flags.numel() == 32 and s.flags.dtype == torch.int32


def _worker_reference_kat(rank, world, port, q):
    """The reference's own multi-GPU known-answer test (tests/modules/losses/test_contrastive_loss_with_temperature.py
    :129-239: global batch 4 split over the ranks, Linear(8,3) / Linear(5,3) encoders under seed 0, GLOBAL backprop):
    mean loss 3.8848, mean image-weight grad 0.0979, mean text-bias grad -1.8151, logit_scale grad 3.6792 — here through
    the oracle's all-gather formulation AND through the LSE-exchange schedule the CUDA path implements."""
    from oracle import clip_oracle as O

    _init(rank, world, port)
    try:
        torch.manual_seed(0)
        image_tensor, text_tensor = torch.randn(4, 8), torch.randn(4, 5)          # fixture order of the reference test
        image_encoder, text_encoder = torch.nn.Linear(8, 3), torch.nn.Linear(5, 3)
        lb = 4 // world
        li, lt = torch.split(image_tensor, lb)[rank], torch.split(text_tensor, lb)[rank]
        scale = torch.tensor(math.log(1 / 0.07), requires_grad=True)
        a, b = image_encoder(li), text_encoder(lt)
        loss = O.contrastive_loss_distributed(a, b, O.clamp_logit_scale(scale), "GLOBAL")[0]
        loss.backward()

        def gathered_mean(x):
            xs = [torch.zeros_like(x) for _ in range(world)]
            dist.all_gather(xs, x.contiguous())
            return torch.stack(xs).mean().item()

        assert abs(gathered_mean(loss.detach().reshape(1)) - 3.8848) < 1e-3
        assert abs(gathered_mean(image_encoder.weight.grad) - 0.0979) < 1e-3
        assert abs(gathered_mean(text_encoder.bias.grad) - (-1.8151)) < 1e-3
        assert abs(gathered_mean(scale.grad.reshape(1)) - 3.6792) < 1e-3
        # the same numbers from the no-gradient-traffic schedule (what engine_loss.contrastive_schedule runs)
        l2, dA, dB, dS = O.contrastive_grads_lse_exchange(a.detach(), b.detach(), scale.detach().clamp(0, math.log(100)), "GLOBAL")
        w_grad = dA.t() @ li              # d loss / d W_image = dA^T x  (Linear backward), per rank
        assert abs(gathered_mean(l2.reshape(1)) - 3.8848) < 1e-3
        assert abs(gathered_mean(w_grad) - 0.0979) < 1e-3
        assert abs(gathered_mean(dB.sum(0)) - (-1.8151)) < 1e-3
        assert abs(gathered_mean(dS.reshape(1)) - 3.6792) < 1e-3
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def test_reference_multi_gpu_loss_known_answers_gloo():
    _run(_worker_reference_kat, world=2)
    _run(_worker_reference_kat, world=1)
