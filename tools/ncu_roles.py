"""Warp-state samples / executed instructions per ROLE of a warp-specialised kernel, with the inlined helpers (mbarrier waits,
conversions) attributed to the role whose code surrounds them in the SASS address order.
    python tools/ncu_roles.py rep.ncu-rep file.cu name:first-last,name:first-last,... [launch]"""
import csv
import io
import subprocess
import sys

rep, main_file, roles_s = sys.argv[1:4]
want = int(sys.argv[4]) if len(sys.argv) > 4 else 0
roles = []
for part in roles_s.split(","):
    n, rng = part.split(":")
    lo, hi = rng.split("-")
    roles.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
launch = -1
fname = None
hdr = None
cur_line = None
sass = []       # (addr, file, line, inst, samples, stalls, text)
seen_files = set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        if fname == main_file and fname in seen_files:
            seen_files = set()
        if fname == main_file:
            launch += 1
        seen_files.add(fname)
        continue
    if r[0] in ("Function Name",):
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or launch != want:
        continue
    if r[0].isdigit():
        cur_line = int(r[0])
        continue
    d = dict(zip(hdr, r))
    addr = d.get("Address", "")
    if not addr.startswith("0x"):
        continue

    def I(k):
        try:
            return int(d.get(k, "0"))
        except ValueError:
            return 0
    stalls = {k[6:]: I(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
    sass.append((int(addr, 16), fname, cur_line, I("Instructions Executed"), I("# Samples"), stalls, d.get("Source", "")))
sass.sort()
label = None
agg = {}
for a, f, ln, inst, smp, st, txt in sass:
    if f == main_file:
        for n, lo, hi in roles:
            if lo <= ln <= hi:
                label = n
                break
    g = agg.setdefault(label or "?", {"inst": 0, "samples": 0, "stalls": {}, "wait_inst": 0, "wait_samp": 0})
    g["inst"] += inst
    g["samples"] += smp
    if "SYNCS" in txt or "NANOSLEEP" in txt:
        g["wait_inst"] += inst
        g["wait_samp"] += smp
    for k, v in st.items():
        g["stalls"][k] = g["stalls"].get(k, 0) + v
ti = sum(g["inst"] for g in agg.values()) or 1
ts = sum(g["samples"] for g in agg.values()) or 1
print("launch %d: %d SASS instructions executed (warp level), %d samples" % (want, ti, ts))
for n, g in agg.items():
    top = sorted(g["stalls"].items(), key=lambda kv: -kv[1])[:6]
    print("  %-10s inst %5.1f%% (mbarrier probes %4.1f%%)  samples %5.1f%% (at mbarrier %4.1f%%)  %s" % (
        n, 100.0 * g["inst"] / ti, 100.0 * g["wait_inst"] / ti, 100.0 * g["samples"] / ts, 100.0 * g["wait_samp"] / ts,
        ", ".join("%s %.0f%%" % (k, 100.0 * v / max(g["samples"], 1)) for k, v in top)))
