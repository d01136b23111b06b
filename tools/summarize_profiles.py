"""Summarise gpurun_out ncu captures into profiles/ (tracked).  usage: python tools/summarize_profiles.py <tag> [out-tag]
  gpurun_out/<tag>_launches.csv : ncu --metrics gpu__time_duration.sum launch list of `bench.py --steps 3 --warmup 3`
  gpurun_out/<tag>_prof.ncu-rep : ncu --set full capture of the top kernels
"""
import collections, csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else tag
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
lines = [f"# ncu summary `{out}` (captured with tools/gpu_round.sh on one B200; cold-cache, serialised launch times: compare shares)\n"]

f = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(f):
    rows = list(csv.reader(open(f)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", "")) / (1000.0 if r[ui] == "ns" else 1.0)
        a = agg.setdefault(r[ki], [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    lines.append(f"## launch list: {sum(a[0] for a in agg.values())} launches, {tot:.0f} us total\n")
    lines.append("| kernel | launches | total us | us/launch | share |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        lines.append(f"| `{k[:90]}` | {n} | {t:.1f} | {t / n:.2f} | {100 * t / tot:.1f}% |")
    lines.append("")

f = os.path.join(G, f"{tag}_prof.ncu-rep")
if os.path.exists(f):
    raw = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, U = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    idx = [(w, H.index(w)) for w in want if w in H]
    lines.append("## `ncu --set full` per-kernel metrics (one row per captured launch)\n")
    lines.append("| " + " | ".join(w.split(".")[0].replace("launch__", "").replace("sm__", "").replace("gpu__", "") for w, _ in idx) + " |")
    lines.append("|" + "---|" * len(idx))
    seen = collections.Counter()
    for r in rows[2:]:
        name = r[idx[0][1]]
        seen[name] += 1
        if seen[name] > 8:
            continue
        lines.append("| " + " | ".join((f"`{r[i][:60]}`" if w == "Kernel Name" else f"{r[i]} {U[i]}") for w, i in idx) + " |")
    lines.append("")
open(os.path.join(P, f"{out}.md"), "w").write("\n".join(lines) + "\n")
print("wrote", os.path.join(P, f"{out}.md"))
