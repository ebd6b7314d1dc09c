#!/usr/bin/env python
"""DRAM traffic and launch times of ONE iteration from an ncu capture -> profiles/r2_iteration_dram.json (read by bench.py).

On the GPU box (one GPU; numbers printed under a profiler are never bench values):

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/iter_dram.csv python bench.py --profile_iters 1

Here:  python profiles/ncu_iteration_traffic.py gpurun_out/iter_dram.csv [--dims 256 128 128] [--precision tf32]

Kernels are grouped into the launch families bench.py reports (forward / data-gradient / weight-gradient convolutions
by matching the launch order against the compiled plan is not possible offline, so the grouping is by kernel name:
`*wgrad*` kernels + their split-K reduce = weight-gradient; the march / halo / conv_tc kernels are split into forward
and data-gradient by their position relative to the loss kernel)."""
import argparse
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--dims", type=int, nargs=3, default=[256, 128, 128])
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_iteration_dram.json"))
    a = ap.parse_args()
    with open(a.csv) as f:
        lines = [l for l in f if not l.startswith("==")]
    launches = collections.OrderedDict()
    for row in csv.DictReader(lines):
        L = launches.setdefault(int(row["ID"]), {"name": re.sub(r"\(.*", "", row["Kernel Name"]), "ns": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(row["Metric Value"].replace(",", ""))
        unit, m = row["Metric Unit"], row["Metric Name"]
        if m == "gpu__time_duration.sum":
            L["ns"] = v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        else:
            b = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            L["rd" if "read" in m else "wr"] = b
    order = list(launches.values())
    loss_at = next((i for i, L in enumerate(order) if "masked_loss" in L["name"]), len(order))
    fam = collections.OrderedDict()
    by_kernel = collections.OrderedDict()
    for i, L in enumerate(order):
        n = L["name"]
        if "wgrad" in n:
            f = "weight-gradient"
        elif "conv_tc" in n or "conv_gather" in n:
            f = "forward" if i < loss_at else "data-gradient"
        else:
            f = "other (BatchNorm / activation / upsample / loss / Adam streams)"
        for key, d in ((f, fam), (n, by_kernel)):
            e = d.setdefault(key, {"launches": 0, "us": 0.0, "dram_bytes": 0.0, "dram_read": 0.0, "dram_write": 0.0})
            e["launches"] += 1
            e["us"] += L["ns"] * 1e-3
            e["dram_read"] += L["rd"]
            e["dram_write"] += L["wr"]
            e["dram_bytes"] += L["rd"] + L["wr"]
    tot = {"launches": len(order), "us": sum(L["ns"] for L in order) * 1e-3,
           "dram_bytes": sum(L["rd"] + L["wr"] for L in order), "dram_read": sum(L["rd"] for L in order),
           "dram_write": sum(L["wr"] for L in order)}
    nvox = a.dims[0] * a.dims[1] * a.dims[2]
    out = {"dims": a.dims, "precision": a.precision,
           "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over one eager iteration "
                     "(bench.py --profile_iters 1), summarised by profiles/ncu_iteration_traffic.py",
           "iteration": tot, "bytes_per_voxel": tot["dram_bytes"] / nvox,
           "ideal_fusion_bytes_per_voxel": 7856 + 512 + 20, "families": fam,
           "kernels": collections.OrderedDict(sorted(by_kernel.items(), key=lambda kv: -kv[1]["us"]))}
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print("%d launches, %.1f us (cold, serialised), DRAM %.2f GB read + %.2f GB written = %.0f B/voxel (ideal fusion %d)"
          % (tot["launches"], tot["us"], tot["dram_read"] / 1e9, tot["dram_write"] / 1e9, out["bytes_per_voxel"],
             out["ideal_fusion_bytes_per_voxel"]))
    for k, e in fam.items():
        print("  %-70s n=%3d %9.1f us %8.2f GB" % (k, e["launches"], e["us"], e["dram_bytes"] / 1e9))


if __name__ == "__main__":
    main()
