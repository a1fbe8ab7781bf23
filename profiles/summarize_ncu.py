#!/usr/bin/env python
"""Text summary of one `ncu --set full --import-source on` capture (a .ncu-rep file): the raw
metrics the roofline discussion uses plus the source lines with the most executed instructions
and the most warp-stall samples.  Usage: python profiles/summarize_ncu.py REPORT.ncu-rep > out.txt"""
import collections
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_atom.sum.pct_of_peak_sustained_elapsed",
    "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.min.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.max.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.min.per_cycle_elapsed", "sm__inst_executed.max.per_cycle_elapsed",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout


def main(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(f"# {rep}: kernel {vals[hdr.index('Kernel Name')]}")
    for i, h in enumerate(hdr):
        if h in RAW or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(vals[i] or 0) > 0.25):
            print(f"{h:88s} {units[i]:10s} {vals[i]}")
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--print-source", "cuda,sass"))))
    cur, hdr, out = None, None, []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) >= 60 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) >= 60 and r[0] != "":
            try:
                ie, sa = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            st = {h: r[i] for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
            top = sorted(((int(v), k) for k, v in st.items() if v not in ("", "-")), reverse=True)[:1]
            out.append((ie, sa, cur, r[0], r[1].strip()[:100], top[0][1] if top and top[0][0] else ""))
    ti, ts = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
    print(f"\n# source lines by warp-stall samples (share of samples / of executed warp instructions)")
    for o in sorted(out, key=lambda o: -o[1])[:16]:
        print(f"  {100 * o[1] / ts:5.1f}% smp {100 * o[0] / ti:5.1f}% inst  {o[2]}:{o[3]}  [{o[5]}]  {o[4]}")
    print(f"\n# source lines by executed warp instructions")
    for o in sorted(out, key=lambda o: -o[0])[:16]:
        print(f"  {100 * o[0] / ti:5.1f}% inst {100 * o[1] / ts:5.1f}% smp  {o[2]}:{o[3]}  {o[4]}")
    print(f"\n# by file")
    byf, bys = collections.Counter(), collections.Counter()
    for o in out:
        byf[o[2]] += o[0]
        bys[o[2]] += o[1]
    for f, c in byf.most_common(10):
        print(f"  {f:45s} {100 * c / ti:5.1f}% inst {100 * bys[f] / ts:5.1f}% smp")


if __name__ == "__main__":
    main(sys.argv[1])
