#!/usr/bin/env python
"""Benchmark of the deep-prior hot path (contract: one JSON line on rank 0).

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --steps K --warmup W    # the reference path on the host cores (oracle port)

A *step* is one optimisation iteration of ``main.py:210-213`` (perturb input, forward, masked loss, backward,
metrics, Adam) on one ``(t,x,y)`` patch of MulResUnet3D at default flags.  Workload at N=1: the patch the
reference's own 3-D number is quoted on — ``(256,128,128)``, inputdepth 64 (``proof_of_concept_3D.ipynb``;
BASELINE.md §1).  N>1: every rank optimises its own patch (patches are independent units — weak scaling, no
data-path collective).

value  = voxel-updates/s over all ranks (iterations/s x 4 194 304 voxels x N), inputs resident in HBM, graph replay.
e2e    = the same metric through the public driver API (``Interpolator.load_data/build_input/optimize``) with
         HOST buffers: H2D of z/img/mask, one D2H of the {loss,snr,pcorr,lr} row per iteration, D2H of out_best.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = ("MulResUnet3D deep-prior optimisation, one (%d,%d,%d) patch per GPU, inputdepth 64, filters 16-256, trilinear "
            "upsample, L1 masked loss, Adam (hyperbolic3d config of BASELINE.json configs[0]; synthetic look-alike volume, "
            "66%% traces removed)")
V100_VOXEL_UPDATES = 1.87e6     # 4 194 304 voxels x 0.445 it/s on a Tesla V100 (notebook output)
FLOP_PER_VOXEL = 391285.5       # fwd+dgrad+wgrad MACs x 2 per voxel per iteration
HBM_BYTES_PER_VOXEL = 7856 + 512 + 20   # SURVEY.md §8d ideal-fusion bytes per voxel per iteration (fp32)


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "which": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update({k: m[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in m})
        p["which"] = "measured"
    except Exception:
        pass
    return p


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def default_args(precision):
    from argparse import Namespace
    return Namespace(datadim="3d", net="multiunet", inputdepth=64, filters=[16, 32, 64, 128, 256],
                     skip=[16, 32, 64, 128], upsample="trilinear", activation="LeakyReLU", last_activation=None,
                     dropout=0., precision=precision, imgchannel=None, loss="mae", epochs=2001, lr=1e-3, lr_factor=.9,
                     lr_thresh=1e-5, lr_patience=100, reduce_lr=False, earlystop_patience=2001, earlystop_min_delta=1.,
                     save_every=None, reg_noise_std=0.03, noise_dist="n", noise_std=.1, data_forgetting_factor=0,
                     filter_noise_with_wavelet=False, lowpass_fs=None, lowpass_fc=None, lowpass_ntaps=7,
                     inittype="xavier", initgain=0.02, netdir=[], savemodel=False, gpu=0, sync_every=1, noise_seed=0,
                     no_cuda_graph=False, start_from_prev=False, gain=40.0)


def synthetic_patch(dims, seed):
    """hyperbolic-event look-alike, 66 % of the traces removed; float64 (t,x,y,1) like data.py"""
    import numpy as np
    rng = np.random.RandomState(seed)
    T, X, Y = dims
    t = np.arange(T)[:, None, None]
    x = np.arange(X)[None, :, None]
    y = np.arange(Y)[None, None, :]
    vol = np.zeros(dims, dtype=np.float64)
    for _ in range(6):
        t0, x0, y0, v = rng.uniform(0.1, 0.8) * T, rng.uniform(0, X), rng.uniform(0, Y), rng.uniform(0.6, 1.5)
        tt = np.sqrt(t0 ** 2 + ((x - x0) ** 2 + (y - y0) ** 2) / v ** 2)
        a = (np.pi * 0.08 * (t - tt)) ** 2
        vol += rng.uniform(0.05, 0.15) * (1 - 2 * a) * np.exp(-a)
    keep = np.ones(X * Y)
    keep[rng.choice(X * Y, int(X * Y * 0.66), replace=False)] = 0
    mask = np.broadcast_to(keep.reshape(1, X, Y), dims).astype(np.float64)
    return (vol * 40.0)[..., None], mask[..., None].copy()


def host_threads():
    """host cores this process may use (the CPU arm uses all of them, whatever OMP_NUM_THREADS says: torchrun exports
    OMP_NUM_THREADS=1 to every rank)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def time_cpu_port(dims, iters, warmup):
    """the oracle port of the reference loop body (main.py:141-213) on ALL host cores -> (voxel-updates/s, s/iter, cores).
    Nothing of the product package is imported here: the initial state_dict comes from oracle.build_state_dict."""
    import torch
    from oracle import net_oracle as O
    torch.set_num_threads(host_threads())
    cores = torch.get_num_threads()
    if cores == 1 and host_threads() > 1:
        raise RuntimeError("CPU baseline would run on one thread of a %d-core host" % host_threads())
    cfg = O.NetConfig()
    sd = O.build_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(1)
    z = torch.randn((1, 64) + dims, generator=g) * 0.1
    img = torch.randn((1, 1) + dims, generator=g)
    mask = (torch.rand((1, 1, 1) + dims[1:], generator=g) > 0.66).float().expand((1, 1) + dims).contiguous()
    st = O.AdamState()
    ts = []
    for i in range(warmup + iters):
        t0 = time.perf_counter()
        O.optimisation_iteration(sd, z, torch.randn(z.shape, generator=g), img, mask, cfg, st, 0.03, "mae", 1e-3)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    nvox = dims[0] * dims[1] * dims[2]
    return nvox / sec, sec, cores


CPU_SAMPLES = [(64, 64, 64), (128, 64, 64), (128, 128, 64), (128, 128, 128), (256, 128, 128)]


def run_reference(a):
    """--impl reference: the reference's CPU path (oracle port; the reference is pure Python / PyTorch, there is nothing
    to compile) on all host cores.  A step = one iteration on the LARGEST sample patch of the same network for which the
    whole --steps/--warmup run stays within ~4 minutes (calibrated with one 64^3 iteration)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    if a.cpu_patch is not None:
        dims = tuple(a.cpu_patch)
    else:
        rate, _, _ = time_cpu_port((64, 64, 64), 1, 1)
        dims = CPU_SAMPLES[0]
        for c in CPU_SAMPLES:
            if (a.steps + a.warmup) * (c[0] * c[1] * c[2]) / rate <= 240.0:
                dims = c
    vps, sec, cores = time_cpu_port(dims, a.steps, a.warmup)
    full = tuple(a.patch) == dims
    line = {"impl": "reference", "metric": "voxel_updates_per_s", "value": vps, "unit": "voxel-updates/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD % tuple(a.patch), "precision": "fp32",
                       "sample": ("the full patch per step" if full else
                                  "a %dx%dx%d patch of the same network and loop body per step (the CPU path needs ~7 s per "
                                  "iteration per 2 M voxels; the metric, voxels x iterations / s, is size-independent to "
                                  "within a few %%: the convolutions dominate at every size)" % dims),
                       "iters_per_s_on_sample": 1.0 / sec, "host_threads": cores},
            "cpu_baseline": {"value": vps, "unit": "voxel-updates/s", "cores": cores, "kind": "port",
                             "sample": "%d timed iteration(s) of one %dx%dx%d patch (oracle port of main.py:141-213 on "
                                       "torch %s CPU ops)" % ((a.steps,) + dims + (__import__("torch").__version__,))},
            "e2e": {"value": vps, "unit": "voxel-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def measure_tf32_peak(dev):
    """dense TF32 GEMM throughput of this GPU, measured the way MEASURED_PEAKS.json measured bf16: torch.matmul
    (cuBLAS) 8192^3 with TF32 operands, best of 10 with CUDA events (burst figure) -> TFLOP/s"""
    import torch
    n = 8192
    x = torch.randn(n, n, device=dev)
    y = torch.randn(n, n, device=dev)
    out = torch.empty(n, n, device=dev)
    old = torch.get_float32_matmul_precision()
    torch.set_float32_matmul_precision("high")
    try:
        best = 1e30
        for i in range(13):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(x, y, out=out)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                best = min(best, e0.elapsed_time(e1))
    finally:
        torch.set_float32_matmul_precision(old)
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def iteration_launches(eng):
    """every pre-marshalled C-ABI call of one iteration of the compiled plan, in issue order:
    (phase, call, op or None, lane)"""
    from deep_prior_interpolation_b200 import engine as E
    calls = [("pack", c, None, 0) for c in eng.pack_calls]
    for phase, lst in (("fwd", eng.fwd_calls), ("loss", [eng.loss_call]), ("bwd", eng.bwd_calls)):
        for c in lst:
            if isinstance(c, E._Wait):
                continue
            for cc in (c.calls if isinstance(c, E._SideCall) else (c,)):
                calls.append((phase, cc, eng.call_op.get(id(cc)), "wgrad" if isinstance(c, E._SideCall) else cc.lane))
    return calls


def time_launches(eng, reps=2, flush=None):
    """CUDA-event duration (us, minimum of `reps`) of every launch of one iteration, each timed alone on the current
    stream (the stream it is launched on).  The working set of an iteration is far larger than L2; `flush` (a buffer
    larger than L2) is rewritten in front of every launch when given."""
    import torch
    from deep_prior_interpolation_b200 import engine as E
    st = E._vp(torch.cuda.current_stream().cuda_stream)
    rows = []
    for phase, c, op, lane in iteration_launches(eng):
        best = 1e30
        for _ in range(reps):
            if flush is not None:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            c(st)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3)
        rows.append((phase, c.name, best, op, lane))
    return rows


CONV_FAMILIES = {"dpi_conv_fwd": "forward", "dpi_conv_fwd_stats": "forward", "dpi_conv_dgrad": "data-gradient",
                 "dpi_conv_wgrad": "weight-gradient"}


def dominant_conv_family(eng, flush):
    """the TIME-DOMINANT family of convolution launches of one iteration (forward / data gradient / weight gradient):
    algorithmic FLOPs (2 x logical Cin x Cout x taps x output voxels per launch) and summed CUDA-event time"""
    from deep_prior_interpolation_b200 import engine as E
    rows = time_launches(eng, reps=2, flush=flush)
    fam = {}
    total_us = 0.0
    for phase, name, us, op, lane in rows:
        total_us += us
        if isinstance(op, E.ConvOp) and name in CONV_FAMILIES:
            f = fam.setdefault(CONV_FAMILIES[name], {"us": 0.0, "flops": 0.0, "launches": 0, "top": (0.0, "")})
            flops = 2.0 * op.y.nvox * op.Cin_l * op.Cout_l * op.taps
            f["us"] += us
            f["flops"] += flops
            f["launches"] += 1
            if us > f["top"][0]:
                f["top"] = (us, "%d->%d k%d %s" % (op.Cin_l, op.Cout_l, round(op.taps ** (1 / 3)) if op.taps > 1 else 1,
                                                    "x".join(str(d) for d in op.x.dims)))
    name = max(fam, key=lambda k: fam[k]["us"])
    return name, fam, total_us, len(rows)


def committed_traffic(dims, precision):
    """DRAM traffic of one iteration from the committed ncu capture (profiles/r2_iteration_dram.json, written by
    profiles/ncu_iteration_traffic.py from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over one eager
    iteration of this workload); None when the capture is for another patch size / precision"""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_iteration_dram.json")) as f:
            t = json.load(f)
        if tuple(t["dims"]) == tuple(dims) and t["precision"] == precision:
            return t
    except Exception:
        pass
    return None


def time_patches_in_flight(dims, precision, ks=(1, 3), steps=20):
    """Secondary workload (BASELINE.json configs[3]: many independent 64^3 patches per GPU): voxel-updates/s with K
    patches in flight, each with its own network / engine / CUDA graph on its own stream, replayed round-robin
    (interpolator._run_patches_in_flight does the same through the public driver).  Device-resident, CUDA events."""
    import torch
    from deep_prior_interpolation_b200.interpolator import Interpolator
    dev = torch.device("cuda", torch.cuda.current_device())
    nvox = dims[0] * dims[1] * dims[2]
    engs, streams, keep = [], [], []
    for i in range(max(ks)):
        args = default_args(precision)
        args.epochs = steps + 8
        img_np, mask_np = synthetic_patch(dims, seed=11 + i)
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
    out = {}
    for k in ks:
        for e in engs[:k]:
            e.reset_loop_state(1e-3, 0)
        torch.cuda.synchronize()
        ms = None
        for rep in range(2):                    # first repetition = warm-up
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s_ in streams[:k]:
                s_.wait_stream(torch.cuda.current_stream())
            n = steps if rep else 3
            for _ in range(n):
                for e, s_ in zip(engs[:k], streams[:k]):
                    with torch.cuda.stream(s_):
                        e.graph.replay()
            for s_ in streams[:k]:
                torch.cuda.current_stream().wait_stream(s_)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
        out[k] = {"ms_per_patch_iteration": ms / k, "voxel_updates_per_s": nvox * k / (ms * 1e-3)}
    return out


def time_lines_2d(precision, steps=200):
    """Secondary workload (BASELINE.json configs[1]): MulResUnet (2-D convs) on the reference's shipped lines example,
    one (170,100) image, the notebook's flags (proof_of_concept_2D.ipynb: datadim 2d, bilinear, default widths; 21.1 it/s on
    a V100).  Device-resident CUDA-graph replay, CUDA events.  The shipped arrays travel under baseline/_ref/ (build());
    without them a synthetic (170,100) stand-in of the same shape is used and named."""
    import numpy as np
    import torch
    from deep_prior_interpolation_b200.interpolator import Interpolator
    dev = torch.device("cuda", torch.cuda.current_device())
    src = os.path.join(ROOT, "baseline", "_ref", "datasets", "lines")
    if os.path.isfile(os.path.join(src, "original.npy")):
        img = np.load(os.path.join(src, "original.npy")).astype(np.float64)
        dec = np.load(os.path.join(src, "random66.npy")).astype(np.float64)
        mask = (dec != 0).astype(np.float64)
        data = "datasets/lines/original.npy + random66.npy (the reference's shipped example)"
    else:
        i3, m3 = synthetic_patch((170, 100, 1), seed=2)
        img, mask = i3[:, :, 0, :] / 40.0, m3[:, :, 0, :]
        data = "synthetic (170,100) stand-in (the shipped lines arrays are not on this box)"
    args = default_args(precision)
    args.datadim, args.upsample, args.gain, args.imgchannel = "2d", "bilinear", 1.0, 1
    args.epochs = steps + 16
    T = Interpolator(args, outpath="/tmp")
    T.load_data({"image": img, "mask": mask, "name": "0"})
    T.build_model()
    T.build_input()
    dims = tuple(img.shape[:2])
    eng = T.net.engine_for(dims, dev, max_iters=args.epochs)
    eng.set_loss("mae")
    eng.set_noise_input(T.input_)
    eng.set_target(T.img_, T.mask_)
    eng.reset_loop_state(1e-3, 0)
    eng.capture(0.03, 0)
    for _ in range(5):
        eng.graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    rec = {"workload": "MulResUnet (2-D), one %dx%d image, inputdepth 64, default widths, bilinear upsample" % dims,
           "data": data, "ms_per_iteration": ms, "it_per_s": 1e3 / ms, "pixel_updates_per_s": dims[0] * dims[1] * 1e3 / ms,
           "launches_per_iteration": eng.launches_per_iteration, "reference_v100_it_per_s": 21.1,
           "reference_source": "proof_of_concept_2D.ipynb:310 (3000 iterations in 2 m 22 s)"}
    T.net.release_engine()
    return rec


def shared_net_record(a, dist, dev, rank, world, dims, steps, barrier):
    """Shared-network mode (BASELINE.json configs[4]): identical weights on every rank, one `dims` patch
    row per rank, local-BN, ONE NCCL all-reduce(sum) of the flat 23.7 MB gradient per iteration between two CUDA graphs
    (rows forward/backward | Adam + bookkeeping), identical fused Adam on every rank.  Times the iteration with and
    without the collective (max over ranks, CUDA events) and checks that the parameters stay bit-identical."""
    import torch
    import deep_prior_interpolation_b200 as dpi
    from deep_prior_interpolation_b200 import utils as u
    from deep_prior_interpolation_b200.distributed import SharedNetTrainer
    args = default_args(a.precision)
    torch.manual_seed(0)
    net = dpi.get_net(args, 1).to(dev)
    u.init_weights(net, "xavier", 0.02)
    eng = net.engine_for(dims, dev, max_iters=4 * steps + 16)
    eng.set_loss("mae")
    img_np, mask_np = synthetic_patch(dims, seed=21 + rank)
    row = eng.new_row(seed=rank)
    img = torch.from_numpy(img_np[..., 0]).float()[None, None].to(dev)
    mask = torch.from_numpy(mask_np[..., 0]).float()[None, None].to(dev)
    eng.row_load(row, (torch.randn((1, 64) + dims) * 0.1).to(dev), img, mask)
    tr = SharedNetTrainer(eng, [row], world, lr=1e-3, sigma=0.03)
    tr.reset()
    tr.capture()
    nvox = dims[0] * dims[1] * dims[2]

    def timed(fn):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    def local_only():          # the same two graphs without the collective = what a patch-sharded rank does
        if tr.graph_all is not None:
            return
        tr.graph_a.replay()
        tr.graph_b.replay()

    ms_local = timed(local_only) if tr.graph_all is None else None
    tr.broadcast_parameters()                         # the local-only replays let the ranks drift apart: re-sync
    tr.reset()
    ms_shared = timed(tr.iteration)
    cs = torch.tensor(tr.param_checksum(), dtype=torch.float64, device=dev)
    same = True
    if dist is not None:
        allc = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allc, cs)
        same = all(torch.equal(c, allc[0]) for c in allc)
    rec = {"workload": "shared-network mode: ONE MulResUnet3D over all rows, one (%d,%d,%d) patch row per GPU, local-BN, "
                       "all-reduce(sum) of the flat gradient between two CUDA graphs, identical fused Adam" % dims,
           "n_gpus": world, "rows": world, "steps": steps, "ms_per_step": ms_shared,
           "voxel_updates_per_s": nvox * world / (ms_shared * 1e-3),
           "ms_per_step_without_allreduce": ms_local,
           "allreduce_overhead": None if ms_local is None else ms_shared / ms_local - 1.0,
           "allreduce_bytes_per_step": int(eng.params.n) * 4,
           "params_bit_identical_across_ranks": bool(same),
           "collective_in_graph": tr.graph_all is not None}
    net.release_engine()
    del tr, eng, net
    torch.cuda.empty_cache()
    return rec


def run_ours(a):
    import torch
    import deep_prior_interpolation_b200 as dpi
    from deep_prior_interpolation_b200 import _lib
    from deep_prior_interpolation_b200.interpolator import Interpolator
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    dims = tuple(a.patch)
    nvox = dims[0] * dims[1] * dims[2]
    args = default_args(a.precision)
    args.epochs = max(a.steps, 8)
    img_np, mask_np = synthetic_patch(dims, seed=1 + rank)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident arm: graph replay of the full iteration -----------------------------
    T = Interpolator(args, outpath="/tmp")
    T.patch_index = rank
    T.load_data({"image": img_np, "mask": mask_np, "name": "0"})
    T.build_model()
    T.build_input()
    eng = T.net.engine_for(dims, dev, max_iters=args.epochs)
    eng.set_loss("mae")
    eng.set_noise_input(T.input_)
    eng.set_target(T.img_, T.mask_)
    eng.reset_loop_state(1e-3, rank)
    eng.capture(0.03, 0)
    for _ in range(max(a.warmup, 3)):
        eng.graph.replay()
    eng.reset_loop_state(1e-3, rank)
    if a.profile_iters > 0:
        # profiling window for `ncu --profile-from-start off`: numbers printed under a profiler are never bench values
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(a.profile_iters):
            eng.iteration(0.03, 0)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profiled_iterations": a.profile_iters, "launches_per_iteration": eng.launches_per_iteration}))
        return
    if a.shared_net:
        rec = shared_net_record(a, dist, dev, rank, world, dims, a.steps, barrier)
        if rank == 0:
            print(json.dumps({"metric": "voxel_updates_per_s", "value": rec["voxel_updates_per_s"], "unit": "voxel-updates/s",
                              "n_gpus": world, "steps": a.steps, "warmup": 3, "ms_per_step": rec["ms_per_step"],
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.precision,
                              "data": "synthetic", "config": rec}))
        if dist is not None:
            dist.destroy_process_group()
        return
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        eng.graph.replay()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / a.steps
    value = nvox * world / (ms_per_step * 1e-3)
    hist = eng.history[:a.steps].cpu().numpy()
    launches = eng.launches_per_iteration * a.steps

    # ---------------- end-to-end arm: public driver API with host buffers ------------------------------------
    args2 = default_args(a.precision)
    args2.epochs = a.steps
    args2.earlystop_patience = a.steps
    T2 = Interpolator(args2, outpath="/tmp")
    T2.patch_index = rank
    T2.net = T.net                      # same compiled plan; a fresh network per patch re-binds it (main.py:286-290)
    T2.load_data({"image": img_np, "mask": mask_np, "name": "0"})
    T2.build_model()                    # per-patch setup (outside the reference's own timer, main.py:209)
    z_host = (torch.randn((1, 64) + dims) * 0.1).pin_memory()
    from deep_prior_interpolation_b200.optim import FusedAdam
    FusedAdam(T2.net)                   # first torch.optim.Optimizer in a process imports torch._dynamo (~2.6 s, once)
    # the end-to-end pass is repeated and the MEDIAN reported: its host part (numpy -> pageable H2D copies, allocator)
    # varies by +-0.15 s from run to run on a fresh box, which is 25 % of a 20-step window
    e2e_all = []
    for _rep in range(3):
        barrier()
        t0 = time.perf_counter()
        T2.load_data({"image": img_np, "mask": mask_np, "name": "0"})      # host numpy -> H2D img, mask
        t1 = time.perf_counter()
        T2.build_input(z_host)                                                 # pinned host z -> H2D
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        T2.optimize()                                                          # K iterations, 1 D2H row each, D2H out_best
        torch.cuda.synchronize()
        e2e_all.append(time.perf_counter() - t0)
        print("e2e split: load_data %.3fs build_input %.3fs optimize %.3fs" % (t1 - t0, t2 - t1, e2e_all[-1] - (t2 - t0)),
              file=sys.stderr)
        with contextlib.redirect_stdout(sys.stderr):
            T2.clean()
        T2.build_model()                # a fresh network for the next pass (outside the timer, as above)
    e2e_s = sorted(e2e_all)[1]
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = nvox * world * a.steps / e2e_s
    h2d = (64 * nvox * 4 + 2 * nvox * 4) / a.steps
    d2h = 32 + nvox * 4 / a.steps

    shared = None
    if world > 1:
        shared = shared_net_record(a, dist, dev, rank, world, (128, 128, 128), 10, barrier)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk = peaks()
    # a seconds-long window of the same graph replays (power / clock behaviour under sustained load; the contract's
    # `value` above stays the K-step number)
    sustained = None
    if world == 1 and a.sustained_s > 0:
        eng.set_noise_input(T.input_)
        eng.set_target(T.img_, T.mask_)
        eng.reset_loop_state(1e-3, rank)
        if eng.graph is None:
            eng.capture(0.03, 0)
        n_s = max(int(a.sustained_s * 1e3 / ms_per_step), a.steps)
        smp = ClockSampler(local)
        torch.cuda.synchronize()
        smp.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_s):
            eng.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ck = smp.stop()
        ms_s = e0.elapsed_time(e1)
        sustained = {"steps": n_s, "seconds": ms_s * 1e-3, "ms_per_step": ms_s / n_s,
                     "voxel_updates_per_s": nvox * n_s / (ms_s * 1e-3), "clocks": ck}
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    fam_name, fams, launches_us, n_launch = dominant_conv_family(eng, flush)
    fam = fams[fam_name]
    k_tflops = fam["flops"] / (fam["us"] * 1e-6) / 1e12
    conv_us = sum(f["us"] for f in fams.values())
    conv_tflops = sum(f["flops"] for f in fams.values()) / (conv_us * 1e-6) / 1e12
    tf32_peak = measure_tf32_peak(dev)
    traffic = committed_traffic(dims, a.precision)
    it_per_s = 1e3 / ms_per_step
    del T2, T, eng
    torch.cuda.empty_cache()
    cpu_dims = tuple(a.cpu_patch) if a.cpu_patch else dims
    cpu_vps, cpu_sec, cores = time_cpu_port(cpu_dims, 1, 1)
    small = time_patches_in_flight((64, 64, 64), a.precision) if world == 1 else None
    lines2d = time_lines_2d(a.precision) if world == 1 else None
    line = {
        "metric": "voxel_updates_per_s", "value": value, "unit": "voxel-updates/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": value / V100_VOXEL_UPDATES, "dtype": "tf32" if a.precision == "tf32" else "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD % dims,
                   "iters_per_s_per_gpu": it_per_s, "voxels_per_patch": nvox, "precision": a.precision,
                   "l2_policy": "per-iteration working set (%.1f GB of activations) far exceeds the 126 MB L2"
                                % (nvox * 2.9e3 / 1e9),
                   "loss_first_last": [float(hist[0, 0]), float(hist[-1, 0])],
                   "vs_baseline_note": "BASELINE.md §1 derived V100 figure (1.87 M voxel-updates/s)"},
        "e2e": {"value": e2e_value, "unit": "voxel-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "seconds": e2e_s, "seconds_all_passes": e2e_all, "passes": "median of 3"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": k_tflops, "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": k_tflops / tf32_peak,
                     "traffic": None if traffic is None else traffic["families"].get(fam_name, {}).get("dram_bytes"),
                     "kernel": "the %d convolution %s launches of one iteration (the time-dominant launch family: %.1f %% of "
                               "the summed launch time; largest single launch %s, %.0f us)"
                               % (fam["launches"], fam_name, 100.0 * fam["us"] / launches_us, fam["top"][1], fam["top"][0]),
                     "kernel_ms": fam["us"] * 1e-3,
                     "algorithmic_flops_per_iteration": fam["flops"],
                     "how": "every launch of the compiled plan timed alone with CUDA events on its stream, L2 flushed "
                            "in front (bench.time_launches); achieved = sum of algorithmic FLOPs (2 x logical Cin x Cout x "
                            "taps x output voxels) / sum of durations over the family; traffic = dram__bytes read+write "
                            "summed over the same launches in the committed ncu capture (profiles/r2_iteration_dram.json)",
                     "families": {k: {"launches": f["launches"], "ms": f["us"] * 1e-3,
                                      "tflops": f["flops"] / (f["us"] * 1e-6) / 1e12} for k, f in fams.items()},
                     "all_convolutions": {"ms": conv_us * 1e-3, "tflops": conv_tflops, "frac": conv_tflops / tf32_peak},
                     "peak_source": "measured in this run: torch.matmul 8192^3 with TF32 operands (cuBLAS), best of 10 "
                                    "(%s bf16 figure of MEASURED_PEAKS.json: %.0f TFLOP/s)" % (pk["which"], pk["bf16_tflops"])},
        "roofline_iteration": {
            "tensor": {"achieved": FLOP_PER_VOXEL * value / world / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                       "frac": FLOP_PER_VOXEL * value / world / 1e12 / tf32_peak},
            "hbm": {"achieved": HBM_BYTES_PER_VOXEL * value / world / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": HBM_BYTES_PER_VOXEL * value / world / 1e9 / pk["hbm_gbs"],
                    "note": "ideal-fusion algorithmic bytes (SURVEY.md §8d), per GPU"},
            "traffic": None if traffic is None else traffic["iteration"]["dram_bytes"],
            "traffic_note": None if traffic is None else
            "dram__bytes_read.sum + dram__bytes_write.sum over the %d launches of one iteration (%s); algorithmic "
            "(ideal-fusion) bytes of the iteration: %.1f GB" % (traffic["iteration"]["launches"], traffic["source"],
                                                                HBM_BYTES_PER_VOXEL * nvox / 1e9),
            "launch_time_sum_ms": launches_us * 1e-3, "launches": n_launch},
        "sustained": sustained,
        "shared_net": shared,
        "lines_2d": lines2d,
        "small_patches": None if small is None else {
            "workload": "BASELINE.json configs[3] patch size: independent 64x64x64 patches on one GPU, device-resident, "
                        "K patches in flight (own network, CUDA graph and stream each; --patches_in_flight)",
            "one_in_flight": small[1], "three_in_flight": small[3],
            "speedup": small[3]["voxel_updates_per_s"] / small[1]["voxel_updates_per_s"]},
        "cpu_baseline": {"value": cpu_vps, "unit": "voxel-updates/s", "cores": cores, "kind": "port",
                         "sample": "1 timed iteration (after 1 warm-up) of one %dx%dx%d patch, %.1f s (oracle port of "
                                   "main.py:141-213 on all host cores)" % (cpu_dims + (cpu_sec,))},
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", type=str, default=os.environ.get("DPI_BENCH_PRECISION", "tf32"), choices=["fp32", "tf32"])
    ap.add_argument("--patch", type=int, nargs=3, default=[256, 128, 128])
    ap.add_argument("--cpu_patch", type=int, nargs=3, default=None,
                    help="sample patch of the CPU arm (default: the GPU arm's own patch for cpu_baseline; --impl reference "
                         "picks the largest sample, up to that patch, that keeps the run within ~4 minutes)")
    ap.add_argument("--sustained_s", type=float, default=10.0,
                    help="N=1: also time a window of about this many seconds of graph replays (0 = off)")
    ap.add_argument("--shared_net", action="store_true",
                    help="N>1: shared-network mode with an NCCL gradient all-reduce per iteration (BASELINE configs[4])")
    ap.add_argument("--profile_iters", type=int, default=0, help="run N eager iterations inside cudaProfilerStart/Stop and exit")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
