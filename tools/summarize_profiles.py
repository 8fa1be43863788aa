"""Turn the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <round-tag> <launches.csv> [<full.ncu-rep> <kernel-label>]

launches.csv: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
               -s 1400 -c 600 --csv --log-file launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`
Writes profiles/<tag>_launch_summary.md (per-kernel launches / time / share / DRAM bytes per launch),
profiles/<tag>_traffic.json (read by bench.py for roofline.traffic) and, with a .ncu-rep, profiles/<tag>_<label>_ncu.md.
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv = sys.argv[1], sys.argv[2]
rep = sys.argv[3] if len(sys.argv) > 3 else None
label = sys.argv[4] if len(sys.argv) > 4 else "conv"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
shutil.copyfile(launches_csv, os.path.join(out_dir, "%s_launches.csv" % tag))

# ---- launch list: per-kernel count / time / share / DRAM traffic (cold-cache, serialised: compare shares)
rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ii, ki, mi, vi, ui = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
per = collections.OrderedDict()          # launch id -> {name, us, rd, wr}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows:
    if r is hdr:
        continue
    d = per.setdefault(r[ii], {"name": r[ki].split("(")[0].replace("void ", "").strip(), "us": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r[vi].replace(",", ""))
    if r[mi] == "gpu__time_duration.sum":
        d["us"] = v / 1000.0 if r[ui] in ("ns", "nsecond") else (v * 1000.0 if r[ui] in ("ms", "msecond") else v)
    elif r[mi] == "dram__bytes_read.sum":
        d["rd"] = v * scale.get(r[ui], 1.0)
    elif r[mi] == "dram__bytes_write.sum":
        d["wr"] = v * scale.get(r[ui], 1.0)
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d["us"]
    a[2] += d["rd"]
    a[3] += d["wr"]
total = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
with open(os.path.join(out_dir, "%s_launch_summary.md" % tag), "w") as f:
    f.write("# %s: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n\n" % tag)
    window = os.environ.get("NCU_WINDOW", "-s 1400 -c 600")          # the launch window the capture script used
    f.write("`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none %s`: "
            "%d launches inside the timed steps, %.1f ms of kernel time.\n" % (window, n, total / 1000))
    f.write("Per-launch times under ncu are cold-cache and serialised: the SHARE column is what compares with bench.py's\n"
            "`roofline.kernels[*].share`; DRAM MB / launch is `dram__bytes_read.sum + dram__bytes_write.sum` averaged over the\n"
            "kernel's launches (compare with the algorithmic bytes of DESIGN.md section 4).\n\n"
            "| kernel | launches | total ms | share | DRAM read MB / launch | DRAM write MB / launch |\n|---|---:|---:|---:|---:|---:|\n")
    for name, (c, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.3f | %.3f | %.1f | %.1f |\n" % (name, c, t / 1000, t / total, rd / c / 1e6, wr / c / 1e6))
traffic = {name: {"launches": c, "dram_bytes_per_launch": (rd + wr) / c, "share": t / total}
           for name, (c, t, rd, wr) in agg.items()}
json.dump({"source": "profiles/%s_launches.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum)" % tag, "kernels": traffic},
          open(os.path.join(out_dir, "%s_traffic.json" % tag), "w"), indent=1)
print("wrote launch summary:", n, "launches")

if rep:
    # ---- full capture of the top kernel: DRAM traffic, throughputs, pipe utilisation per launch
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    idx = {w: h.index(w) for w in want if w in h}
    with open(os.path.join(out_dir, "%s_%s_ncu.md" % (tag, label)), "w") as f:
        f.write("# %s: `ncu --set full --clock-control none --import-source on` of %s launches inside `bench.py --steps 2 --warmup 3`\n\n" % (tag, label))
        f.write("traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch (compare with the algorithmic bytes in DESIGN.md).\n\n")
        f.write("| # | " + " | ".join(w.split(".")[0].replace("__", ":") for w in idx) + " |\n|---|" + "---:|" * len(idx) + "\n")
        for i, r in enumerate(rr[2:]):
            f.write("| %d | " % i + " | ".join("%s %s" % (r[j], units[j]) for j in idx.values()) + " |\n")
    print("wrote ncu summary:", len(rr) - 2, "kernels")
