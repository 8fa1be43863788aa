"""Summaries of an .ncu-rep (read here, no GPU): headline metrics per launch and, with --roles, warp-state samples and executed
instructions per role of a warp-specialised kernel (source line ranges given as name:first-last,...).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--roles producer:300-380,mma:381-450,...] [--top 25]"""
import argparse
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
           "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_xu.sum", "lts__t_sector_hit_rate.pct",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        res.append(d)
    return hdr, res


def source(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    launches = []
    cur = None
    hdr = None
    fname = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            if fname and fname.startswith("conv_") and (cur is None or fname in cur["files"]):
                cur = {"files": set(), "lines": []}
                launches.append(cur)
            cur["files"].add(fname)
            continue
        if hdr and r[0].isdigit() and cur is not None:
            d = dict(zip(hdr, r))

            def I(k):
                try:
                    return int(d.get(k, "0"))
                except ValueError:
                    return 0
            stalls = {k: I(k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
            cur["lines"].append((fname, int(r[0]), r[1], I("Instructions Executed"), I("# Samples"), stalls))
    return launches


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--roles", default="")
    ap.add_argument("--top", type=int, default=0)
    ap.add_argument("--file", default="")
    a = ap.parse_args()
    hdr, res = raw(a.rep)
    for i, d in enumerate(res):
        print("launch %d: %s" % (i, d.get("Kernel Name", "")[:90]))
        for m in METRICS:
            if m in d:
                print("   %-80s %s" % (m, d[m]))
    if not a.roles and not a.top:
        return
    launches = source(a.rep)
    roles = []
    for part in a.roles.split(","):
        if part:
            n, rng = part.split(":")
            lo, hi = rng.split("-")
            roles.append((n, int(lo), int(hi)))
    for li, L in enumerate(launches):
        lines = L["lines"]
        tot_i = sum(x[3] for x in lines) or 1
        tot_s = sum(x[4] for x in lines) or 1
        print("== launch %d: %d instructions, %d samples" % (li, tot_i, tot_s))
        main_file = a.file or sorted(L["files"])[0]
        for n, lo, hi in roles:
            sel = [x for x in lines if x[0] == main_file and lo <= x[1] <= hi]
            si, ss = sum(x[3] for x in sel), sum(x[4] for x in sel)
            st = {}
            for x in sel:
                for k, v in x[5].items():
                    st[k] = st.get(k, 0) + v
            top = sorted(st.items(), key=lambda kv: -kv[1])[:5]
            print("   %-10s inst %5.1f%%  samples %5.1f%%   %s" % (n, 100.0 * si / tot_i, 100.0 * ss / tot_s,
                                                               ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(ss, 1)) for k, v in top)))
        other = [x for x in lines if x[0] != main_file]
        print("   %-10s inst %5.1f%%  samples %5.1f%%" % ("helpers", 100.0 * sum(x[3] for x in other) / tot_i, 100.0 * sum(x[4] for x in other) / tot_s))
        if a.top:
            for x in sorted(lines, key=lambda x: -x[4])[:a.top]:
                st = sorted(x[5].items(), key=lambda kv: -kv[1])[:3]
                print("      %-14s %4d  inst %5.2f%% samp %5.2f%%  %-70s %s" % (x[0], x[1], 100.0 * x[3] / tot_i, 100.0 * x[4] / tot_s, x[2].strip()[:70],
                                                                           ",".join("%s:%d" % (k[6:], v) for k, v in st)))


if __name__ == "__main__":
    main()
