"""Turns an `ncu --set full` report of scripts/ncu_kernels.py into profiles/<tag>_ncu_kernels.{csv,json}: per launch the
duration, DRAM bytes, achieved DRAM GB/s, tensor-pipe and MUFU (xu) utilisation, issue-slot utilisation, registers.
Runs here (no GPU needed): python scripts/ncu_summarize.py gpurun_out/r2_kernels.ncu-rep r2"""
import csv
import io
import json
import subprocess
import sys

rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg"]
have = [w for w in WANT if w in col]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def scale(metric, v):
    u = units[col[metric]]
    if v is None:
        return None
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0,
            "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
    return v * mult[u] if u in mult else v


GEMM_ORDER = ["gemm_qkv_fwd", "gemm_fc1_act", "gemm_fc2_dgrad_dact", "gemm_fc1_wgrad"]   # launch order in ncu_kernels.py


def key_of(kname, seen):
    """Stable key from the kernel name (the backward is one fused kernel or a dQ + dK/dV pair, depending on the build)."""
    if "gemm_kernel" in kname:
        k = GEMM_ORDER[seen["gemm"]] if seen["gemm"] < len(GEMM_ORDER) else f"gemm{seen['gemm']}"
        seen["gemm"] += 1
        return k
    if "attn_fwd" in kname:
        return "attn_fwd"
    if "attn_bwd_fused" in kname:
        return "attn_bwd_fused"
    if "attn_bwd" in kname:
        seen["bwd"] += 1
        return "attn_bwd_dq" if seen["bwd"] == 1 else "attn_bwd_dkdv"
    if "add_ln_fwd" in kname:
        return "add_ln_fwd"
    if "ln_bwd" in kname:
        return "ln_bwd"
    return None
M, d, ff, B, S, H = 201728, 768, 3072, 1024, 197, 12
ALGO = {  # algorithmic bytes / flops of the launch
    "gemm_qkv_fwd": (M * d * 2 + 3 * d * d * 2 + M * 3 * d * 2, 2.0 * M * 3 * d * d),
    "gemm_fc1_act": (M * d * 2 + ff * d * 2 + 2 * M * ff * 2, 2.0 * M * ff * d),
    "gemm_fc2_dgrad_dact": (M * d * 2 + ff * d * 2 + 2 * M * ff * 2, 2.0 * M * ff * d),
    "gemm_fc1_wgrad": (M * ff * 2 + M * d * 2 + ff * d * 4, 2.0 * M * ff * d),
    "attn_fwd": (M * 3 * d * 2 + M * d * 2, 4.0 * S * S * 64 * H * B),
    "attn_bwd_dq": (M * 3 * d * 2 + 2 * M * d * 2 + M * d * 2, 4.0 * S * S * 64 * H * B * 1.5),
    "attn_bwd_dkdv": (M * 3 * d * 2 + M * d * 2 + 2 * M * d * 2, 4.0 * S * S * 64 * H * B * 2.0),
    "attn_bwd_fused": (M * 3 * d * 2 + 2 * M * d * 2 + M * 3 * d * 2, 4.0 * S * S * 64 * H * B * 2.5),
    "add_ln_fwd": (M * d * 12, 0.0),
    "ln_bwd": (M * d * 16, 0.0),
}
out = {"source": f"profiles/{tag}_ncu_kernels.csv (ncu --set full --clock-control none, scripts/ncu_kernels.py, B=1024 shapes)",
       "kernels": {}}
lines = [["launch", "kernel"] + have + ["algorithmic_bytes", "algorithmic_flops", "dram_GBps", "TFLOPs"]]
kcol = col.get("Kernel Name")
seen = {"gemm": 0, "bwd": 0}
for i, r in enumerate(data):
    name = key_of(r[kcol], seen) or f"launch{i}"
    vals = {m: scale(m, num(r[col[m]])) for m in have}
    ab, af = ALGO.get(name, (None, None))
    t = vals.get("gpu__time_duration.sum")
    rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
    gbps = (rd + wr) / t / 1e9 if (t and rd is not None and wr is not None) else None
    tf = af / t / 1e12 if (t and af) else None
    lines.append([name, r[kcol][:70]] + [vals[m] for m in have] + [ab, af, gbps, tf])
    out["kernels"][name] = {"name": r[kcol].split("(")[0][:90], "shape": f"B={B} S={S} H={H} d={d} ff={ff} (M={M})",
                            "duration_s": t, "dram_bytes_read": rd, "dram_bytes_write": wr, "algorithmic_bytes": ab,
                            "algorithmic_flops": af, "dram_GBps": gbps, "TFLOPs": tf,
                            **{m: vals[m] for m in have if m not in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")}}
with open(f"profiles/{tag}_ncu_kernels.csv", "w", newline="") as f:
    csv.writer(f).writerows(lines)
json.dump(out, open(f"profiles/{tag}_ncu_kernels.json", "w"), indent=1)
for l in lines:
    print(l[0], l[1][:40], [f"{x:.4g}" if isinstance(x, float) else x for x in l[2:]])
