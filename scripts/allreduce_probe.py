"""Times the parameter-gradient all-reduce ALONE (NCCL, one process per GPU under torchrun): the 600 MB fp32 flat
gradient of the CLIP ViT-B/16 step as two buffers (text 254 MB, image 345 MB), and the same payload in bf16.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/allreduce_probe.py
Writes gpurun_out/r2_allreduce_probe_N.log (rank 0)."""
import os

import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W, rank = dist.get_world_size(), dist.get_rank()
lines = []


def t(bufs, name, iters=10):
    for _ in range(3):
        for b in bufs:
            dist.all_reduce(b)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        works = [dist.all_reduce(b, async_op=True) for b in bufs]
        for w in works:
            w.wait()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nbytes = sum(b.numel() * b.element_size() for b in bufs)
    if rank == 0:
        lines.append(f"N={W} {name}: {ms:.3f} ms per step-worth of gradients ({nbytes / 1e6:.0f} MB, algbw {nbytes / ms / 1e6:.0f} GB/s, "
                     f"busbw {nbytes / ms / 1e6 * 2 * (W - 1) / W:.0f} GB/s)")


n_img, n_txt = 86_192_640, 63_428_096
t([torch.zeros(n_txt, device=dev), torch.zeros(n_img, device=dev)], "fp32, two flat buffers")
t([torch.zeros(n_txt + n_img, device=dev)], "fp32, one buffer")
t([torch.zeros(n_txt, device=dev, dtype=torch.bfloat16), torch.zeros(n_img, device=dev, dtype=torch.bfloat16)], "bf16, two flat buffers")
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    open(f"gpurun_out/r2_allreduce_probe_{W}.log", "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
dist.destroy_process_group()
