#!/usr/bin/env python
"""Top stall locations of the first kernel in an ncu report (needs `--import-source on` / `--set full` at capture time):

    python profiles/ncu_hot.py gpurun_out/x.ncu-rep [--top 25] [--context 6]

Prints the headline metrics (duration, DRAM bytes, tensor-pipe activity) and the instructions with the most warp-stall
samples, each with its dominant stall reasons and a few preceding SASS lines."""
import argparse
import csv
import io
import subprocess


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_active.avg",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__inst_executed.sum", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    for r in rows[2:3]:
        print(r[hdr.index("Kernel Name")][:110])
        for i, h in enumerate(hdr):
            if h in want:
                print("   %-70s %-10s %s" % (h, units[i], r[i]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--context", type=int, default=6)
    a = ap.parse_args()
    raw_metrics(a.rep)
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[ia].startswith("0x"):
            if r and r[0] == "Kernel Name":
                break
            continue
        data.append(r)
    tot = sum(int(r[isamp]) for r in data) or 1
    print("%d instructions, %d samples" % (len(data), tot))
    order = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:a.top]
    for i in order:
        r = data[i]
        st = sorted(((int(r[c]), h) for c, h in stall_cols), reverse=True)[:2]
        print("%6d %5.1f%% executed %9s  %s   %s" % (int(r[isamp]), 100.0 * int(r[isamp]) / tot, r[iex], r[isrc].strip()[:70],
                                                      ", ".join("%s %d" % (h, n) for n, h in st if n)))
        for rr in data[max(0, i - a.context):i]:
            print("            . %6s %9s  %s" % (rr[isamp], rr[iex], rr[isrc].strip()[:90]))


if __name__ == "__main__":
    main()
