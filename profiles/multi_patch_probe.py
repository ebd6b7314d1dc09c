#!/usr/bin/env python
"""How much does keeping K independent patches in flight on ONE GPU buy?  (BASELINE configs[3]: 1470 patches of 64^3)

Each patch has its own network / engine / captured graph; the K graphs are replayed round-robin on K streams.

    python profiles/multi_patch_probe.py --patch 64 64 64 --k 1 2 3 4 --steps 30
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--patch", type=int, nargs=3, default=[64, 64, 64])
    ap.add_argument("--k", type=int, nargs="+", default=[1, 2, 3, 4])
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    import torch
    import bench
    from deep_prior_interpolation_b200.interpolator import Interpolator
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dims = tuple(a.patch)
    nvox = dims[0] * dims[1] * dims[2]
    kmax = max(a.k)
    engs, streams, keep = [], [], []
    for i in range(kmax):
        args = bench.default_args("tf32")
        args.epochs = a.steps + 8
        img_np, mask_np = bench.synthetic_patch(dims, seed=1 + i)
        T = Interpolator(args, outpath="/tmp")
        T.patch_index = i
        T.load_data({"image": img_np, "mask": mask_np, "name": str(i)})
        T.build_model()
        T.build_input()
        eng = T.net.engine_for(dims, dev, max_iters=args.epochs)
        eng.set_loss("mae")
        eng.set_noise_input(T.input_)
        eng.set_target(T.img_, T.mask_)
        eng.reset_loop_state(1e-3, i)
        eng.capture(0.03, 0)
        engs.append(eng)
        streams.append(torch.cuda.Stream(dev))
        keep.append(T)
    torch.cuda.synchronize()
    for k in a.k:
        for e in engs[:k]:
            e.reset_loop_state(1e-3, 0)
        torch.cuda.synchronize()
        for rep in range(2):        # first repetition = warm-up
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in streams[:k]:
                s.wait_stream(torch.cuda.current_stream())
            for _ in range(a.steps if rep else 3):
                for e, s in zip(engs[:k], streams[:k]):
                    with torch.cuda.stream(s):
                        e.graph.replay()
            for s in streams[:k]:
                torch.cuda.current_stream().wait_stream(s)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        print("patch %s  K=%d in flight: %.3f ms per round, %.3f ms per patch-iteration, %.1f M voxel-updates/s"
              % (dims, k, ms, ms / k, nvox * k / ms / 1e3), flush=True)


if __name__ == "__main__":
    main()
