"""2-rank tests (one rank per GPU when 2 GPUs are visible; both ranks time-share cuda:0 otherwise — never skipped on
a GPU box): the distributed contrastive loss — peers' embeddings read in place over
NVLink by the GEMMs' TMA producers, LSE-exchange backward — against the oracle's restatement of the reference's
all_gather-with-backprop formulation (utils/distributed.py:28-58), for the three BackpropTypes; and the data-parallel
trainer against a single process running the concatenated global batch.

Mirrors tests/modules/losses/test_contrastive_loss_with_temperature.py:129-239 and tests/utils/test_distributed.py
(spawned workers, file-free 127.0.0.1 rendezvous, nccl)."""
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _init(rank, world, port):
    """One process per GPU over NCCL when the box has >= `world` GPUs.  On a 1-GPU lease the same two ranks SHARE
    cuda:0 (time-sliced), rendezvous over gloo, and still exchange embeddings through CUDA-IPC peer mappings and the
    release/acquire flag kernel — the code path under test is identical (NCCL itself refuses two ranks on one GPU;
    it only carries the parameter-gradient all-reduce of the trainer test, which gloo performs here)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    n = torch.cuda.device_count()
    local = rank if n >= world else 0
    torch.cuda.set_device(local)
    if n >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    return torch.device("cuda", local)


def _worker_loss(rank, world, port, q):
    from multimodal_b200.modules.losses.contrastive_loss_with_temperature import contrastive_loss_with_temperature
    from multimodal_b200.utils.distributed import BackpropType
    from oracle import clip_oracle as O

    dev = _init(rank, world, port)
    try:
        torch.manual_seed(100 + rank)
        B, E = 128, 256
        a0 = O.normalize(torch.randn(B, E, device=dev))
        b0 = O.normalize(torch.randn(B, E, device=dev))
        row_mask = torch.rand(B, device=dev) < (0.8 if rank == 0 else 0.4)   # different kept-row counts per rank
        row_mask[0] = True
        for mode in (BackpropType.GLOBAL, BackpropType.LOCAL, BackpropType.NONE):
            for eps, mask in ((0.0, None), (0.1, None), (0.1, row_mask)):
                a, b = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
                s = torch.tensor(math.log(1 / 0.07), device=dev, requires_grad=True)
                res = contrastive_loss_with_temperature(a, b, s, backprop_type=mode, mask=mask,
                                                        cross_entropy_kwargs={"label_smoothing": eps})
                res.loss.backward()
                ar, br = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
                sr = torch.tensor(math.log(1 / 0.07), device=dev, requires_grad=True)
                ref = O.contrastive_loss_distributed(ar, br, sr, mode.name, eps, mask)
                ref[0].backward()
                assert abs(res.loss.item() - ref[0].item()) < 3e-3, (mode, eps, res.loss.item(), ref[0].item())
                assert (res.logits_a - ref[1]).abs().max().item() < 3e-2     # bf16 embeddings x T=14.3
                assert (res.logits_b - ref[2]).abs().max().item() < 3e-2
                for got, want, name in ((a.grad, ar.grad, "dA"), (b.grad, br.grad, "dB")):
                    rel = ((got - want).abs().max() / want.abs().max()).item()
                    assert rel < 3e-2, (mode, eps, name, rel)
                assert abs(s.grad.item() - sr.grad.item()) < 5e-3 * max(1.0, abs(sr.grad.item())), (mode, eps)
                # the FUSED schedule (what the trainer runs: per-peer similarity GEMMs whose epilogues keep the logits in
                # registers) against the materialising one just checked
                from multimodal_b200.engine_loss import contrastive_schedule
                f = contrastive_schedule(a0, b0, s.detach().reshape(1), eps, mode, False, world, rank, mask)
                assert abs(f[0].item() - res.loss.item()) < 2e-5 * max(1.0, abs(res.loss.item())), (mode, eps)
                for got, want, name in ((f[5], a.grad, "dA"), (f[6], b.grad, "dB")):
                    rel = ((got - want).abs().max() / want.abs().max().clamp_min(1e-20)).item()
                    assert rel < 1e-2, ("fused", mode, eps, name, rel)
                assert abs(f[7].item() - s.grad.item()) < 1e-4 * max(1.0, abs(s.grad.item())), ("fused", mode, eps)
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def _small(dev, sd=None):
    from multimodal_b200.models.clip.image_encoder import CLIPViTEncoder
    from multimodal_b200.models.clip.model import CLIP
    from multimodal_b200.models.clip.text_encoder import CLIPTextEncoder

    m = CLIP(CLIPViTEncoder(64, 16, 64, 128, 2, 2),
             CLIPTextEncoder(embedding_dim=64, vocab_size=512, width=128, dim_feedforward=512, heads=2, layers=2))
    if sd is not None:
        m.load_state_dict(sd)
    return m.to(dev).train()


def _grads_of_step(trainer, img, txt):
    import multimodal_b200.ops as ops

    orig, seen = ops.adamw_step, {}

    def spy(p, g, *a, **kw):
        seen[g.data_ptr()] = g.clone()
        return orig(p, g, *a, **kw)

    ops.adamw_step = spy
    try:
        loss = trainer.step(img, txt)
    finally:
        ops.adamw_step = orig
    return loss, seen[trainer.img.store.g.data_ptr()], seen[trainer.txt.store.g.data_ptr()]


def _worker_trainer(rank, world, port, q):
    from multimodal_b200.modules.losses.contrastive_loss_with_temperature import ContrastiveLossWithTemperature
    from multimodal_b200.train import ContrastiveTrainer
    from oracle import clip_oracle as O

    dev = _init(rank, world, port)
    try:
        torch.manual_seed(0)
        m = _small(dev)
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        Bl = 64
        img_all, txt_all = O.synthetic_batch(Bl * world, image_size=64, vocab=512, device=dev)
        tr = ContrastiveTrainer(m, ContrastiveLossWithTemperature().to(dev), lr=0.0, weight_decay=0.0)
        assert tr.world == world
        loss, gi, gt = _grads_of_step(tr, img_all[rank * Bl:(rank + 1) * Bl].contiguous(),
                                      txt_all[rank * Bl:(rank + 1) * Bl].contiguous())
        gi, gt = gi / world, gt / world      # AdamW applies grad_scale = 1/world to the all-reduced sum
        lsum = loss.detach().clone()
        dist.all_reduce(lsum)
        if rank == 0:
            # single-process reference: world forced to 1 on the concatenated global batch
            import multimodal_b200.engine_loss as el
            import multimodal_b200.train as trn
            saved = el._dist_state
            el._dist_state = trn._dist_state = lambda: (1, 0)
            try:
                m1 = _small(dev, sd)
                tr1 = ContrastiveTrainer(m1, ContrastiveLossWithTemperature().to(dev), lr=0.0, weight_decay=0.0)
                loss1, gi1, gt1 = _grads_of_step(tr1, img_all, txt_all)
            finally:
                el._dist_state = trn._dist_state = saved
            assert abs(lsum.item() / world - loss1.item()) < 2e-3, (lsum.item() / world, loss1.item())
            for got, want, name in ((gi, gi1, "image"), (gt, gt1, "text")):
                rel = ((got - want).abs().max() / want.abs().max()).item()
                assert rel < 2e-2, (name, rel)
        q.put((rank, "ok"))
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    mode = "one rank per GPU, nccl" if torch.cuda.device_count() >= world else f"{world} ranks sharing cuda:0, gloo rendezvous"
    print(f"[distributed test] {world} ranks, {torch.cuda.device_count()} visible GPU(s): {mode}")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 200)
    procs = [ctx.Process(target=fn, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    ok = all(p.exitcode == 0 for p in procs)
    res = sorted(q.get(timeout=5) for _ in range(world)) if ok else []
    assert ok and res == [(r, "ok") for r in range(world)]


def test_distributed_contrastive_loss_2gpu():
    _run(_worker_loss)


def test_data_parallel_trainer_matches_single_process_2gpu():
    _run(_worker_trainer)
