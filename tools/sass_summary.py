"""SASS evidence of the Blackwell-native code paths (B200_PROFILING.md, "What proves a Blackwell-native kernel"):
    python tools/sass_summary.py [out.md]
Disassembles the built library with cuobjdump and counts, per kernel, the mnemonics that tcgen05.mma (UTC*MMA), tcgen05.ld/st
(LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP), mbarrier (SYNCS), setmaxnreg (USETMAXREG), st.async / cluster traffic and
the legacy tensor path (HMMA: must be zero) compile to."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "styletts2_lite_b200", "lib", "libst2_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_summary.md")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
PAT = collections.OrderedDict([("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
                               ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("USETMAXREG", r"\bUSETMAXREG"),
                               ("UTCBAR/commit", r"\bUTCBAR"), ("ST.ASYNC", r"\bSTAS|ST\.ASYNC"), ("HMMA (legacy)", r"\bHMMA"),
                               ("HGMMA (Hopper)", r"\b[HQI]GMMA")])
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        cur = per.setdefault(name, collections.Counter())
        continue
    if cur is None or "/*" not in line:
        continue
    for k, pat in PAT.items():
        if re.search(pat, line):
            cur[k] += 1
agg = collections.OrderedDict()
for name, c in per.items():
    base = re.sub(r"<.*", "", name)
    a = agg.setdefault(base, [0, collections.Counter()])
    a[0] += 1
    a[1].update(c)
tot = collections.Counter()
with open(out, "w") as f:
    f.write("# SASS mnemonic counts of `styletts2_lite_b200/lib/libst2_b200.so` (cuobjdump -sass, sm_100a)\n\n")
    f.write("Written by `tools/sass_summary.py`.  Per kernel family (template instantiations summed): `UTC*MMA` = tcgen05.mma, "
            "`LDTM` / `STTM` = tcgen05.ld / st, `UTMALDG` / `UTMASTG` / `UBLKCP` = TMA (cp.async.bulk[.tensor]), `SYNCS` = mbarrier, "
            "`USETMAXREG` = setmaxnreg, `UTCBAR` = tcgen05.commit.  `HMMA` (mma.sync / wmma) and `*GMMA` (wgmma) must be zero: "
            "no legacy tensor path is compiled in.\n\n")
    keys = list(PAT.keys())
    f.write("| kernel | instantiations | " + " | ".join(keys) + " |\n|---|---:|" + "---:|" * len(keys) + "\n")
    for base, (n, c) in agg.items():
        if not any(c[k] for k in keys):
            continue
        tot.update(c)
        f.write("| `%s` | %d | " % (base, n) + " | ".join(str(c[k]) for k in keys) + " |\n")
    f.write("| **all kernels** | %d | " % sum(a[0] for a in agg.values()) + " | ".join("**%d**" % tot[k] for k in keys) + " |\n")
print("wrote", out, dict(tot))
