#!/usr/bin/env python
"""Per-launch timing of one iteration with CUDA events (no profiler), each launch annotated with its layer.

    python profiles/op_profile.py [--patch 256 128 128] [--precision tf32] [--reps 3] > profiles/rN_op_profile.txt

Every pre-marshalled C-ABI call of the compiled plan (engine.Engine) is replayed `reps` times between two events
with an L2 flush in front; the minimum is reported.  Convolutions also print their MMA-issue floor
(SURVEY.md App. B shapes; 44/48/64/128 clk per kind::tf32 MMA for N <= 32/64/128/256, profiles/r1_probe_umma_issue_rate.txt)
and their ideal HBM time at the measured copy bandwidth, so the gap of each launch to its own bound is visible.
"""
import argparse
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patch", type=int, nargs=3, default=[256, 128, 128])
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--min_us", type=float, default=0.0)
    a = ap.parse_args()
    import torch
    import bench
    import deep_prior_interpolation_b200 as dpi
    from deep_prior_interpolation_b200 import engine as E
    from deep_prior_interpolation_b200.interpolator import Interpolator
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dims = tuple(a.patch)
    args = bench.default_args(a.precision)
    args.epochs = 8
    img_np, mask_np = bench.synthetic_patch(dims, seed=1)
    T = Interpolator(args, outpath="/tmp")
    T.patch_index = 0
    T.load_data({"image": img_np, "mask": mask_np, "name": "0"})
    T.build_model()
    T.build_input()
    eng = T.net.engine_for(dims, dev, max_iters=8)
    eng.set_loss("mae")
    eng.set_noise_input(T.input_)
    eng.set_target(T.img_, T.mask_)
    eng.reset_loop_state(1e-3, 0)
    for _ in range(2):
        eng.iteration(0.03, 0)
    torch.cuda.synchronize()

    def tag(op):
        if isinstance(op, E.ConvOp):
            g = op.geom
            return "conv %d->%d k%d s%d in %dx%dx%d (Cp %d->%d)" % (op.Cin_l, op.Cout_l, g.kh, g.stride if hasattr(g, "stride") else 0,
                                                                    op.x.dims[0], op.x.dims[1], op.x.dims[2], op.x.C, op.y.C)
        if isinstance(op, E.BnActOp):
            return "bnact C%d vox %d bn=%d" % (op.x.C, op.x.nvox, op.bn is not None)
        if isinstance(op, E.AddActOp):
            return "addact C%d vox %d bn=%d parts=%d" % (op.C, op.nvox, op.bn is not None, len(op.qs))
        if isinstance(op, E.UpsampleOp):
            return "upsample C%d -> vox %d" % (op.x.C, op.out.nvox)
        return type(op).__name__

    calls = []
    for c in eng.pack_calls:
        calls.append(("pack", c, None))
    lane_of = {}
    for op in eng.ops:
        for c in op.emit_fwd():
            if not isinstance(c, E._Wait):
                calls.append(("fwd", c, op))
                lane_of[id(c)] = getattr(op, "lane", 0)
    calls.append(("loss", eng.loss_call, None))
    for op in reversed(eng.ops):
        for c in op.emit_bwd():
            if isinstance(c, E._Wait):
                continue
            for cc in (c.calls if isinstance(c, E._SideCall) else (c,)):
                calls.append(("bwd", cc, op))
                lane_of[id(cc)] = "wgrad" if isinstance(c, E._SideCall) else getattr(op, "lane", 0)
    calls.append(("bwd", eng.bwd_calls[-1], None))      # batched gradient un-pack

    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    st = E._vp(torch.cuda.current_stream().cuda_stream)
    clk = 1.965e9

    def mma_clk(n):
        return 44 if n <= 32 else 48 if n <= 64 else 64 if n <= 128 else 128

    tot = {}
    lanes = {}
    rows = []
    for phase, c, op in calls:
        best = 1e30
        for _ in range(a.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            c(st)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3)
        extra = ""
        if isinstance(op, E.ConvOp) and c.name in ("dpi_conv_fwd", "dpi_conv_fwd_stats", "dpi_conv_dgrad", "dpi_conv_wgrad"):
            taps = op.taps
            vo, vi = op.y.nvox, op.x.nvox
            ci, co = op.x.C, op.y.C
            if c.name in ("dpi_conv_fwd", "dpi_conv_fwd_stats"):
                M, N, K = vo, co, taps * math.ceil(ci / 8) * 8
            elif c.name == "dpi_conv_dgrad":
                M, N, K = vi, ci, taps * math.ceil(co / 8) * 8 * (vo / vi if vo < vi else 1)
            else:
                M, N, K = taps * ci, co, vo
            nt = min(256, math.ceil(N / 16) * 16)
            floor = math.ceil(M / 128) * math.ceil(N / nt) * math.ceil(K / 8) * mma_clk(nt) / 148 / clk * 1e6
            hbm = 4.0 * (vi * ci + vo * co) / 6.54e12 * 1e6
            extra = "  mma_floor %.0f us  hbm %.0f us" % (floor, hbm)
        key = phase + ":" + c.name
        tot[key] = tot.get(key, 0.0) + best
        lk = (phase if phase in ("fwd", "bwd") else "fwd" if phase == "pack" else "bwd", lane_of.get(id(c), 0))
        lanes[lk] = lanes.get(lk, 0.0) + best
        rows.append((phase, c.name, best, tag(op) if op is not None else "", extra))
    total = sum(tot.values())
    print("patch %s precision %s: %d launches, %.1f us summed (each launch timed alone after an L2 flush)"
          % (dims, a.precision, len(rows), total))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print("%-34s %10.1f us %5.1f%%" % (k, v, 100 * v / total))
    print()
    print("per launch lane (engine.Engine._run: lane 0 = main chain, 1 = shortcut branches, 2+ = ResPath of a level, "
          "wgrad = weight-gradient lane); lanes of one phase run concurrently:")
    for k in sorted(lanes, key=lambda kv: (kv[0], str(kv[1]))):
        print("  %-4s lane %-6s %10.1f us" % (k[0], k[1], lanes[k]))
    print()
    for i, (phase, name, us, tg, extra) in enumerate(rows):
        if us >= a.min_us:
            print("%4d %-4s %-24s %9.1f us  %s%s" % (i, phase, name, us, tg, extra))


if __name__ == "__main__":
    main()
