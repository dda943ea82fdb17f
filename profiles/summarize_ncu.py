#!/usr/bin/env python
"""Turns the raw ncu outputs a gpurun call brought back into the small text summaries kept in profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches_r1.csv          > profiles/r1_launches.txt
    python profiles/summarize_ncu.py full     gpurun_out/prof_r1.ncu-rep          > profiles/r1_field_tc_ncu.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split("(")[0].replace("void ", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("# per-kernel totals of `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: read SHARES)")
    print("%-72s %6s %12s %12s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s %6d %12.1f %12.2f %7.3f" % (k[:72], n, t / 1e3, t / 1e3 / n, t / tot))
    print("%-72s %6d %12.1f" % ("TOTAL", sum(a[0] for a in agg.values()), tot / 1e3))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print("# `ncu --set full --clock-control none --import-source on`, selected raw metrics per captured launch")
    for r in rows[2:]:
        print("\n== %s   grid %s block %s" % (r[ki].split("(")[0], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("  %-82s %14s %s" % (m, r[i], units[i]))


def _rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def _gb(val, unit):
    """ncu prints dram bytes in a unit of its choosing per column (byte, Kbyte, Mbyte, Gbyte)."""
    f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(val.replace(",", "")) * f


def traffic(field_rep, train_rep):
    """profiles/traffic.json: what bench.py reports as roofline.traffic (it cannot run ncu inside a timed run).

        python profiles/summarize_ncu.py traffic gpurun_out/prof_r2c_field.ncu-rep gpurun_out/prof_r2c_train.ncu-rep > profiles/traffic.json

    field_tc_kernel: dram__bytes_read.sum + dram__bytes_write.sum averaged over the captured launches of the INFERENCE
    kernel (one step = coarse/fine x fg/bg).  train_step: the same summed over every captured launch of one training
    step's kernels (training forward, dgrad chain, wgrad, heads)."""
    import json
    import os
    out = {}
    hdr, units, rows = _rows(field_rep)
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    # template arguments <BG, CLUSTER, TRAIN, PREC>: the fast inference kernel is <., 1, 0, 0> (ncu prints bools as 0 / 1 or as
    # (bool)0 / (bool)1 depending on the page)
    norm = lambda k: k.replace("(bool)", "").replace("(int)", "").replace("false", "0").replace("true", "1")
    vals = [_gb(r[ri], units[ri]) + _gb(r[wi], units[wi]) for r in rows if "field_tc_kernel" in r[ki] and ", 0, 0>" in norm(r[ki])]
    out["field_tc_kernel"] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals),
                              "capture": "ncu --set full --clock-control none, " + os.path.basename(field_rep)}
    hdr, units, rows = _rows(train_rep)
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    per = collections.OrderedDict()
    for r in rows:
        k = r[ki].split("(")[0].replace("void ", "")
        per[k] = per.get(k, 0.0) + _gb(r[ri], units[ri]) + _gb(r[wi], units[wi])
    out["train_step"] = {"dram_bytes_per_launch": sum(per.values()), "per_kernel": per, "launches": len(rows),
                         "capture": "ncu --set full --clock-control none, one training step (both levels), " + os.path.basename(train_rep),
                         "note": "bytes per STEP (sum over the step's field kernels)"}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
