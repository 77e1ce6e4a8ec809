"""Aggregates the ncu source page (per-instruction stall samples) of one kernel into segments that end at a
synchronisation landmark (mbarrier try_wait / arrive, tcgen05.ld, tcgen05.commit, BAR): shows where the warps of a
warp-specialised kernel spend their time.   python scripts/ncu_segments.py <report.ncu-rep> <kernel regex> [min %] [n-th matching launch]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]   # one block per matching launch
nth = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if len(starts) >= 2 and all(rows[starts[i]] == rows[starts[i + 1]] for i in range(0, len(starts) - 1, 2)):
    starts = starts[::2] + [len(rows)]   # ncu prints every launch twice (SASS view, source view): keep the first
    rows_end = {starts[i]: starts[i + 1] for i in range(len(starts) - 1)}
rows = rows[starts[nth]:(starts[nth + 1] if nth + 1 < len(starts) else len(rows))]
print(rows[0][1][:110])
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
seen, data = set(), []
for r in rows[hi + 1:]:
    try:
        a = int(r[ia], 16)
    except (ValueError, IndexError):
        continue
    if a in seen:
        continue
    seen.add(a)
    st = {h[i]: int(r[i] or 0) for i in stall_cols}
    data.append((a, int(r[isamp] or 0), int(r[iex] or 0), r[isrc], st))
data.sort()
tot = sum(d[1] for d in data) or 1
totex = sum(d[2] for d in data) or 1
print(f"samples {tot}  warp-instructions {totex}")
seg_s = seg_e = 0
seg_st = {}
base = data[0][0]
for a, s, e, src, st in data:
    seg_s += s
    seg_e += e
    for k, v in st.items():
        seg_st[k] = seg_st.get(k, 0) + v
    if any(k in src for k in ("SYNCS.PHASECHK", "LDTM", "EXIT", "BAR.SYNC", "UTCBAR", "SYNCS.ARRIVE", "WARPSYNC")):
        if seg_s > minp / 100 * tot:
            top = sorted(seg_st.items(), key=lambda kv: -kv[1])[:3]
            tops = " ".join(f"{k[6:]}={100 * v / max(1, sum(seg_st.values())):.0f}%" for k, v in top)
            print(f"{a - base:6x}  samples {100 * seg_s / tot:5.1f}%  instr {100 * seg_e / totex:5.1f}%  [{tops}]  -> {src[:64]}")
        seg_s = seg_e = 0
        seg_st = {}
