#!/usr/bin/env python
"""End-to-end demonstration of BASELINE.json configs[3] at reduced size, through the public command line:

    python profiles/run_config4_demo.py [--shape 192 128 128] [--epochs 200] [--patches_in_flight 0]
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 ... profiles/run_config4_demo.py   (patch-sharded)

Writes a synthetic (t,x,y) volume of hyperbolic Ricker events with 70 % of the traces removed at random (NaN traces,
data.py:53-54), runs ``deep_prior_interpolation_b200.interpolator.main`` with 64^3 patches at 50 % stride, reassembles
with ``reconstruct_patches`` and prints ONE JSON line: wall time of the whole job (patch extraction, per-patch set-up,
optimisation, result files), patch-iterations/s, voxel-updates/s, and the SNR of the reassembled volume on all traces
and on the removed traces only.  (Wall-clock numbers of a whole job; bench.py holds the timed-kernel numbers.)
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def volume(shape, seed):
    rng = np.random.RandomState(seed)
    t = np.arange(shape[0])[:, None, None]
    x = np.arange(shape[1])[None, :, None]
    y = np.arange(shape[2])[None, None, :]
    vol = np.zeros(shape)
    for _ in range(8):
        t0, x0, y0 = rng.uniform(8, shape[0] * 0.9), rng.uniform(0, shape[1]), rng.uniform(0, shape[2])
        v = rng.uniform(0.6, 1.6)
        tt = np.sqrt(t0 ** 2 + ((x - x0) ** 2 + (y - y0) ** 2) / v ** 2)
        a = (np.pi * 0.10 * (t - tt)) ** 2
        vol += rng.uniform(0.05, 0.15) * (1 - 2 * a) * np.exp(-a)
    return vol


def snr(out, ref):
    return float(10 * np.log10(np.sum(ref ** 2) / np.sum((ref - out) ** 2)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs=3, default=[192, 128, 128])
    ap.add_argument("--epochs", type=int, default=200)
    ap.add_argument("--patches_in_flight", type=int, default=0)
    ap.add_argument("--rate", type=float, default=0.7)
    ap.add_argument("--workdir", type=str, default=None)
    a = ap.parse_args()
    import torch
    from deep_prior_interpolation_b200 import interpolator, data as D, utils as u
    from deep_prior_interpolation_b200.parameter import parse_arguments
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    work = a.workdir or os.path.join(tempfile.gettempdir(), "dpi_config4_demo")
    os.makedirs(work, exist_ok=True)
    os.chdir(work)
    shape = tuple(a.shape)
    # rank 0 synthesises the volume (tens of seconds of numpy at 1000x256x256); the other ranks read its files
    vol = mask = None
    if rank == 0:
        vol = volume(shape, 4)
        np.random.seed(4)
        mask = u.build_mask(vol, a.rate)
        dec = vol.copy()
        dec[mask == 0] = np.nan
        np.save("original.npy", vol)
        np.save("decimated.npy", dec)
        del dec
    if world > 1:
        dist.barrier()
    argv = ["--imgdir", work, "--imgname", "original.npy", "--maskname", "decimated.npy", "--datadim", "3d", "--gain", "40",
            "--upsample", "linear", "--patch_shape", "64", "64", "64", "--patch_stride", "32", "32", "32",
            "--epochs", str(a.epochs), "--precision", "tf32", "--outdir", "demo", "--sync_every", "25",
            "--patches_in_flight", str(a.patches_in_flight), "--gpu", str(int(os.environ.get("LOCAL_RANK", 0)))]
    if rank == 0:
        for f in os.listdir(os.path.join(work, "results", "demo")) if os.path.isdir(os.path.join(work, "results", "demo")) else []:
            os.remove(os.path.join(work, "results", "demo", f))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    interpolator.main(argv)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    if rank == 0:
        args = parse_arguments(argv)
        rec = D.reconstruct_patches(args)
        ref = vol[:rec.shape[0], :rec.shape[1], :rec.shape[2]]
        m = mask[:rec.shape[0], :rec.shape[1], :rec.shape[2]]
        n_patches = len([f for f in os.listdir(os.path.join("results", "demo")) if f.endswith("_run.npy")])
        k = interpolator.patches_in_flight(args, (64, 64, 64), (n_patches + world - 1) // world)
        print(json.dumps({
            "workload": "synthetic %dx%dx%d volume, %.0f%% traces removed, 64^3 patches at 50%% stride, %d epochs per patch"
                        % (shape + (100 * a.rate, a.epochs)),
            "n_gpus": world, "patches": n_patches, "patches_in_flight_per_gpu": k, "wall_s": wall,
            "patch_iterations_per_s": n_patches * a.epochs / wall,
            "voxel_updates_per_s": n_patches * a.epochs * 64 ** 3 / wall,
            "snr_db_all_traces": snr(rec, ref), "snr_db_removed_traces": snr(rec[m == 0], ref[m == 0]),
            "snr_db_input": snr(ref * m, ref)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
