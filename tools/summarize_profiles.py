"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <round-tag> <launches.csv> <full.ncu-rep>
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel count / time / share (cold-cache, serialised: compare shares)
rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
total = 0.0
n = 0
for r in rows:
    if r is hdr or r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000.0 if r[ui] in ("ns", "nsecond") else (v * 1000.0 if r[ui] in ("ms", "msecond") else v)   # -> us
    name = r[ki].split("(")[0].replace("void ", "").strip()
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    total += v
    n += 1
with open(os.path.join(out_dir, "%s_launch_summary.md" % tag), "w") as f:
    f.write("# %s: ncu launch list of `python bench.py --steps 2 --warmup 3` (gpu__time_duration.sum, --clock-control none)\n\n" % tag)
    f.write("%d launches captured (-s 1400 -c 600: inside the timed/profiled steps), %.1f ms of kernel time.\n" % (n, total / 1000))
    f.write("Per-launch times under ncu are cold-cache and serialised: the SHARE column is what compares with bench.py's\n"
            "`roofline.kernels[*].share`.\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.3f | %.3f |\n" % (name, c, t / 1000, t / total))
print("wrote launch summary:", n, "launches")

# ---- full capture of the top kernel: DRAM traffic, throughputs, pipe utilisation per launch
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
idx = {w: h.index(w) for w in want if w in h}
with open(os.path.join(out_dir, "%s_conv_fused_ncu.md" % tag), "w") as f:
    f.write("# %s: `ncu --set full --clock-control none` of conv_fused_kernel launches inside `bench.py --steps 2 --warmup 3`\n\n" % tag)
    f.write("traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch (compare with the algorithmic bytes in DESIGN.md).\n\n")
    f.write("| # | " + " | ".join(w.split(".")[0].replace("__", ":") for w in idx) + " |\n|---|" + "---:|" * len(idx) + "\n")
    for i, r in enumerate(rr[2:]):
        f.write("| %d | " % i + " | ".join("%s %s" % (r[j], units[j]) for j in idx.values()) + " |\n")
print("wrote ncu summary:", len(rr) - 2, "kernels")
