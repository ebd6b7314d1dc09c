#!/usr/bin/env python
"""ms per iteration (CUDA-graph replay, device-resident) for one patch size - the quick A/B tool behind the numbers in
DESIGN.md:  python profiles/quick_iter_time.py [--patch 256 128 128] [--steps 20]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patch", type=int, nargs=3, default=[256, 128, 128])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--precision", default="tf32")
    a = ap.parse_args()
    import torch
    import bench
    from deep_prior_interpolation_b200.interpolator import Interpolator
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dims = tuple(a.patch)
    args = bench.default_args(a.precision)
    args.epochs = a.steps + 8
    img_np, mask_np = bench.synthetic_patch(dims, seed=1)
    T = Interpolator(args, outpath="/tmp")
    T.load_data({"image": img_np, "mask": mask_np, "name": "0"})
    T.build_model()
    T.build_input()
    eng = T.net.engine_for(dims, dev, max_iters=args.epochs)
    eng.set_loss("mae")
    eng.set_noise_input(T.input_)
    eng.set_target(T.img_, T.mask_)
    eng.reset_loop_state(1e-3, 0)
    eng.capture(0.03, 0)
    for _ in range(3):
        eng.graph.replay()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            eng.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / a.steps)
    print("patch %s %s lib=%s: %.3f ms per iteration" % (dims, a.precision, os.path.basename(os.environ.get("DPI_B200_LIB", "default")), best))


if __name__ == "__main__":
    main()
