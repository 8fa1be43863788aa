"""Where a warp-specialised kernel waits: samples at every mbarrier wait / TMEM load / barrier in SASS address order.
    python tools/ncu_waits.py rep.ncu-rep [launch]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; k = -1; ker = []
for r in rows:
    if not r: continue
    if r[0] == "Kernel Name": k += 1; ker.append([]); continue
    if r[0] == "Address": hdr = r; continue
    if hdr and r[0].startswith("0x"):
        d = dict(zip(hdr, r)); ker[k].append((int(r[0], 16), r[1].strip(), int(d["# Samples"]), int(d["Instructions Executed"])))
K = ker[want]; base = K[0][0]; ts = sum(x[2] for x in K) or 1
print("launch %d: %d samples, %d warp instructions" % (want, ts, sum(x[3] for x in K)))
for a, t, s, i in K:
    if any(w in t for w in ("SYNCS.PHASECHK", "NANOSLEEP", "LDTM", "UTMALDG", "BAR.SYNC", "USETMAXREG", "UTCHMMA")) and (s > ts * 0.004 or "USETMAXREG" in t):
        print("%6x  samples %5.2f%%  executed %9d  %s" % (a - base, 100.0 * s / ts, i, t[:80]))
