"""Driver with the reference's ``main.py`` surface: ``Interpolator`` and ``main()``.

The per-patch loop, file formats (``args.txt``, ``<name>_run.npy``, ``<name>_model.pth``, ``<name>_output<iter>.npy``)
and method names follow ``main.py:18-297``.  The loop body (``main.py:141-193,210-217``) — perturb input, forward,
masked loss, backward, metrics, best-output tracking, Adam — runs as ONE replayed CUDA graph of hand-written
kernels (``engine.Engine.iteration``); the host only reads back ``{loss, snr, pcorr, lr}`` rows.

Patch-level data parallelism: launched under ``torchrun`` every rank takes the patches
``i % WORLD_SIZE == RANK``; patches are independent, so there is no collective on this path.
"""
from __future__ import annotations

import os
import warnings
from time import time

import numpy as np
import torch

from . import utils as u
from .architectures import get_net
from .data import extract_patches, history_alias
from .optim import FusedAdam
from .parameter import net_args_are_same, parse_arguments

__all__ = ["Interpolator", "main", "patches_in_flight"]


class Interpolator:
    def __init__(self, args, outpath):
        self.args = args
        if args.gpu is None or not torch.cuda.is_available():
            raise RuntimeError("deep_prior_interpolation_b200 needs a CUDA device (--gpu); there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = torch.cuda.FloatTensor
        self.outpath = outpath
        if args.loss not in ("mae", "mse"):
            raise ValueError("loss must be mae or mse")
        self.elapsed = None
        self.iiter = 0
        self.iter_to_be_saved = list(range(0, args.epochs, int(args.save_every))) \
            if args.save_every is not None else [0]
        self.loss_min = None
        self.outchannel = args.imgchannel
        self.history = u.History(args.epochs)
        self.image_name = None
        self.img = self.img_ = self.mask = self.mask_ = None
        self.out_best = None
        self.zfill = u.ten_digit(args.epochs)
        self.input_type = "noise3d" if args.datadim == "3d" else "noise"
        self.input_ = None
        self.input_list = []
        self.net = None
        self.num_params = None
        self.optimizer = None
        self.patch_index = 0

    # -- per-patch setup (main.py:59-139) ------------------------------------------------------------------
    def build_input(self, z_host: torch.Tensor = None):
        """z ~ noise_dist * noise_std, drawn on the CPU generator like main.py:59-64 (or supplied, already scaled)"""
        a = self.args
        data_shape = self.img.shape[:-1]
        if z_host is not None:
            assert tuple(z_host.shape) == (1, a.inputdepth) + tuple(data_shape)
            self.input_ = z_host.to(self.device, non_blocking=True)
        else:
            self.input_ = u.get_noise((1, a.inputdepth) + tuple(data_shape), a.noise_dist).to(self.device)
            self.input_ *= a.noise_std
        if a.filter_noise_with_wavelet:
            # main.py:66-72: the noise takes the bandwidth of the source wavelet stored next to the data
            self.input_ = u.fir_time(self.input_, np.load(os.path.join(a.imgdir, "wavelet.npy")))
        if a.lowpass_fs and a.lowpass_fc:
            # main.py:74-84: 4th-order Butterworth response, realised as --lowpass_ntaps FIR taps
            print("filtering the input tensor with a low pass Butterworth...")
            taps = u.lowpass_butterworth_taps(fc=a.lowpass_fc, fs=a.lowpass_fs, ntaps=a.lowpass_ntaps, order=4,
                                              nfft=2 ** u.nextpow2(self.input_.shape[2]))
            self.input_ = u.fir_time(self.input_, taps)
        self.add_data_ = self.add_data_weight = None
        if a.data_forgetting_factor != 0:
            # main.py:86-97: decimated data normalised to the std of the input noise; the repetition along the input
            # depth (data_.repeat(...)[:, :inputdepth]) is a channel modulo inside dpi_add_data_dev, but the std is
            # taken over the repeated-and-cropped tensor like the reference does
            data_ = self.img_ * self.mask_
            num_rep = int(np.ceil(self.input_.shape[1] / data_.shape[1]))
            rep = data_.repeat([1, num_rep] + [1] * len(data_shape))[:, :a.inputdepth]
            self.add_data_ = data_ * (torch.std(self.input_) / torch.std(rep))
            self.add_data_weight = np.logspace(0, -4, a.data_forgetting_factor)
            del rep
        print("The input shape is %s" % str(tuple(self.input_.shape)))

    def build_model(self, netpath: str = None):
        a = self.args
        if self.outchannel is None:
            self.outchannel = self.img_.shape[1]
        if a.netdir is not None and len(a.netdir) != 0:
            _args = u.read_args(os.path.join("./results", *netpath.split("/")[:-1], "args.txt"))
            if not hasattr(_args, "precision"):
                _args.precision = a.precision
            assert net_args_are_same(a, _args)
            self.net = get_net(_args, self.outchannel).to(self.device)
            self.net.load_state_dict(torch.load(os.path.join("./results", netpath), map_location=self.device))
            print("Network loaded from %s" % os.path.join("./results", netpath))
        else:
            old = self.net
            self.net = get_net(a, self.outchannel).to(self.device)
            u.init_weights(self.net, a.inittype, a.initgain)
            if old is not None and old._engine is not None and old._engine.rebind(self.net):
                object.__setattr__(self.net, "_engine", old._engine)   # reuse the compiled plan + CUDA graph
                old.release_engine()
        self.num_params = sum(int(np.prod(list(p.size()))) for p in self.net.parameters())

    def load_data(self, data):
        self.image_name = data["name"]
        self.img = data["image"]
        self.mask = data["mask"]
        if self.mask.shape != self.img.shape:
            raise ValueError("The loaded mask shape has to be", self.img.shape)
        sha = tuple(range(self.img.ndim))
        re_sha = sha[-1:] + sha[:-1]
        self.img_ = u.np_to_torch(np.transpose(self.img, re_sha), bc_add=False).unsqueeze(0).float().to(self.device)
        self.mask_ = u.np_to_torch(np.transpose(self.mask, re_sha), bc_add=False).unsqueeze(0).float().to(self.device)
        return torch.std(self.img_ * self.mask_).item()

    # -- the hot loop (main.py:141-220) ------------------------------------------------------------------------
    def _np_out(self, out_: torch.Tensor) -> np.ndarray:
        return u.torch_to_np(out_, True) if out_.ndim > 4 else u.torch_to_np(out_, False)[0].transpose((1, 2, 0))

    def optimize(self):
        """``optimize`` of main.py:195-220: the whole loop of one patch on the current stream"""
        self._opt_begin()
        torch.cuda.synchronize(self.device)
        self._opt_start_clock()
        done = False
        while not done:
            self._opt_launch()
            done = self._opt_collect()
        self._opt_end()

    # The loop is split into begin / launch / collect / end so that several patches can be kept in flight on one GPU,
    # each on its own stream (``_run_patches_in_flight``); ``optimize`` above is the one-patch composition of the four.
    def _opt_begin(self):
        a = self.args
        print("starting optimization with ADAM...")
        t_setup = time()
        had_engine = self.net._engine is not None
        eng = self.net.engine_for(self.input_.shape[2:], self.device, max_iters=a.epochs)
        tt = [time()]
        eng.set_loss(a.loss)
        eng.set_noise_input(self.input_)
        eng.set_target(self.img_, self.mask_)
        eng.set_data_forgetting(getattr(self, "add_data_", None), getattr(self, "add_data_weight", None))
        tt.append(time())
        self.optimizer = FusedAdam(self.net, lr=a.lr)
        tt.append(time())
        scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, mode="min", factor=a.lr_factor,
                                                               threshold=a.lr_thresh, patience=a.lr_patience)
        stopper = u.EarlyStopping(patience=a.earlystop_patience, min_delta=a.earlystop_min_delta, percentage=True)
        sigma = float(a.reg_noise_std) if a.reg_noise_std > 0 else 0.0
        seed = int(getattr(a, "noise_seed", 0)) * 1000003 + self.patch_index
        eng.reset_loop_state(a.lr, seed)
        use_graph = not getattr(a, "no_cuda_graph", False)
        if use_graph and (eng.graph is None or getattr(eng, "_graph_sigma", None) != sigma):
            eng.capture(sigma, 0)
        # host decisions that depend on every iteration's loss force a per-iteration read-back
        sync_every = max(1, int(getattr(a, "sync_every", 1)))
        if a.reduce_lr or a.earlystop_patience < a.epochs:
            sync_every = 1
        save_at = sorted(i for i in self.iter_to_be_saved if i != 0)
        tt.append(time())
        if os.environ.get("DPI_TIMING"):
            print("optimize setup: had_engine=%s engine_for %.3f set_inputs %.3f adam %.3f sched+capture %.3f"
                  % (had_engine, tt[0] - t_setup, tt[1] - tt[0], tt[2] - tt[1], tt[3] - tt[2]))
        self._opt = {"eng": eng, "scheduler": scheduler, "stopper": stopper, "sigma": sigma, "use_graph": use_graph,
                     "sync_every": sync_every, "save_at": save_at, "j": 0, "n": 0, "t_setup": t_setup,
                     "start": time()}

    def _opt_start_clock(self):
        self._opt["start"] = time()

    def _opt_launch(self):
        """enqueue the next chunk of iterations on the current stream (no host<->device sync)"""
        o, a = self._opt, self.args
        j = o["j"]
        n = min(o["sync_every"], a.epochs - j)
        if j < a.data_forgetting_factor:
            n = 1                                        # the network input of these iterations is recorded (main.py:155)
        nxt = [s for s in o["save_at"] if j <= s < j + n]
        if nxt:
            n = nxt[0] - j + 1
        eng = o["eng"]
        for _ in range(n):
            if o["use_graph"]:
                eng.graph.replay()
            else:
                eng.iteration(o["sigma"], 0)
        o["n"] = n

    def _opt_collect(self) -> bool:
        """read the chunk's {loss, snr, pcorr, lr} rows back and run the host-side bookkeeping; True when finished"""
        o, a = self._opt, self.args
        eng, j, n = o["eng"], o["j"], o["n"]
        stop = False
        rows = eng.history[j:j + n].cpu().numpy()        # the only host<->device sync of the loop
        for r in range(n):
            l, s, p, lr = (float(v) for v in rows[r])
            self.history.append((l, s, p))
            self.history.lr.append(self.optimizer.param_groups[0]["lr"])
            if not o.get("quiet"):
                print(self.history.log_message(self.iiter), "\r", end="")
            if self.iiter == 0 or l <= self.loss_min:
                self.loss_min = l
            if self.iiter < a.data_forgetting_factor:
                self.input_list.append(u.torch_to_np(eng.network_input_nchw(), True))
            if self.iiter in o["save_at"]:
                np.save(os.path.join(self.outpath, self.image_name.split(".")[0] + "_output%s.npy"
                                     % str(self.iiter).zfill(self.zfill)), self._np_out(eng.output_nchw()))
            self.iiter += 1
            if a.reduce_lr:
                o["scheduler"].step(l)
                new_lr = self.optimizer.param_groups[0]["lr"]
                if new_lr != lr:
                    eng.set_lr(new_lr)
            if o["stopper"].step(l):
                stop = True
                break
        o["j"] = j + n
        return stop or o["j"] >= a.epochs

    def _opt_end(self):
        o = self._opt
        torch.cuda.current_stream(self.device).synchronize()
        self.elapsed = time() - o["start"]
        t_loop = time()
        self.out_best = self._np_out(o["eng"].output_nchw(best=True))
        if os.environ.get("DPI_TIMING"):
            print("optimize split: setup %.3fs loop %.3fs out_best %.3fs"
                  % (o["start"] - o["t_setup"], self.elapsed, time() - t_loop))
        print(u.sec2time(self.elapsed))
        self._opt = None

    def save_result(self):
        with history_alias() as History:
            hist = self.history
            if type(hist) is not History:
                # pickle under the reference's class path so the file loads in either code base (main.py:226-235)
                h2 = History.__new__(History)
                h2.__dict__.update(hist.__dict__)
                hist = h2
            np.save(os.path.join(self.outpath, self.image_name + "_run.npy"), {
                "device": u.get_gpu_name(),
                "elapsed": u.sec2time(self.elapsed),
                "outpath": self.outpath,
                "history": hist,
                "mask": self.mask,
                "image": self.img,
                "output": self.out_best,
                "noise": self.input_list,
            })
        if self.args.savemodel and self.net is not None:
            torch.save({k: v.detach().clone() for k, v in self.net.state_dict().items()},
                       os.path.join(self.outpath, self.image_name + "_model.pth"))

    def clean(self):
        self.iiter = 0
        print("Finished patch %s" % self.image_name)
        self.loss_min = None
        self.history = u.History(self.args.epochs)
        self.input_list = []        # (the reference never empties it, so every patch's *_run.npy also carries the
                                    # recorded inputs of all earlier patches: main.py:234,241-251)


def _rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def patches_in_flight(args, patch_shape, n_patches: int) -> int:
    """How many independent patches to optimise concurrently on this GPU (``--patches_in_flight``, 0 = automatic).

    Patches are independent problems (own network, noise, Adam state; main.py:274-295), so several of them can share
    the GPU, each replaying its own CUDA graph on its own stream.  Small patches leave most of the 148 SMs idle in the
    lower U-Net levels and every launch of their dependent chain costs latency rather than throughput: on a B200,
    three 64^3 patches in flight finish 1.44x more iterations per second than one (3.69 vs 5.31 ms per
    patch-iteration), two 128^3 patches 1.06x; one (256,128,128) patch fills the GPU by itself."""
    k = int(getattr(args, "patches_in_flight", 1))
    if args.start_from_prev:
        return 1                                    # sequential by definition (main.py:286)
    if k <= 0:
        nvox = int(np.prod([int(v) for v in patch_shape]))
        k = 3 if nvox <= 80 ** 3 else (2 if nvox <= 128 ** 3 else 1)
    return max(1, min(k, n_patches))


def _setup_patch(T: "Interpolator", i: int, patch, args) -> bool:
    """per-patch setup of main.py:277-290; False when the patch is all zeros and was written out without optimising"""
    T.patch_index = i
    print("\nThe data shape is %s, " % str(patch["image"].shape), end="")
    std = T.load_data(patch)
    print("the std of coarse data is %.2e" % std)
    if np.isclose(std, 0., atol=1e-12):
        print("skipping...")
        T.out_best = T.img * T.mask
        T.elapsed = 0.
        return False
    if T.net is None or not args.start_from_prev:
        if args.netdir is not None and len(args.netdir) != 0:
            T.build_model(netpath=args.netdir[i])
        else:
            T.build_model()
    T.build_input()
    return True


def _run_patches_in_flight(args, outpath, mine, k: int) -> None:
    """Round-robin scheduler over ``k`` slots.  A slot is an Interpolator (its own network, engine, captured graph) on
    its own stream; the host collects the finished chunk of one slot and immediately launches that slot's next chunk
    while the other k-1 slots keep the GPU busy, so read-backs, host bookkeeping and the per-patch setup of the next
    patch overlap with compute.  Per-patch results do not depend on k: setup runs in patch order (the same sequence of
    RNG draws as one-at-a-time) and nothing is shared between slots."""
    dev = torch.device("cuda", torch.cuda.current_device())
    slots = [Interpolator(args, outpath) for _ in range(k)]
    streams = [torch.cuda.Stream(dev) for _ in range(k)]
    queue = iter(mine)

    def feed(T: "Interpolator") -> bool:
        for i, patch in queue:
            if _setup_patch(T, i, patch, args):
                T._opt_begin()
                T._opt["quiet"] = True
                T._opt_start_clock()
                T._opt_launch()
                return True
            T.save_result()
            T.clean()
        return False

    active = []
    for T, st in zip(slots, streams):
        st.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(st):
            active.append(feed(T))
    while any(active):
        for n, (T, st) in enumerate(zip(slots, streams)):
            if not active[n]:
                continue
            with torch.cuda.stream(st):
                if T._opt_collect():
                    T._opt_end()
                    T.save_result()
                    T.clean()
                    active[n] = feed(T)
                else:
                    T._opt_launch()
    torch.cuda.synchronize(dev)


def _run_shared_net(args, outpath, patches) -> None:
    """``--shared_net`` (BASELINE config 5; not in the reference, whose loop builds one network per patch,
    main.py:274-295): ONE network is optimised over all patches.  Batch rows = patches, rank r holds the rows
    ``p % WORLD_SIZE == r``; every iteration each rank runs forward/backward on its rows, the flat gradient is summed
    over ranks with one NCCL all-reduce and every rank applies the identical fused Adam step
    (``distributed.SharedNetTrainer``).  Per-row results are written as the usual ``<name>_run.npy`` files by the rank
    that holds the row, so ``reconstruct_patches`` works unchanged; rank 0 writes ``shared_model.pth``."""
    from . import distributed as D
    a = args
    rank, world = _rank_world()
    dev = torch.device("cuda", torch.cuda.current_device())
    if a.data_forgetting_factor != 0 or a.start_from_prev:
        raise NotImplementedError("--shared_net does not combine with --data_forgetting_factor / --start_from_prev")
    if world > 1:
        D.init_process_group(dev)
    # all-zero patches are written out without optimising, like main.py:281-284; they are not batch rows
    live = [i for i, p in enumerate(patches) if not np.isclose(float(np.std(p["image"] * p["mask"])), 0., atol=1e-12)]
    holders = []
    for i, patch in enumerate(patches):
        if i % world != rank:
            continue
        T = Interpolator(a, outpath)
        T.patch_index = i
        T.load_data(patch)
        if i not in live:
            print("patch %s is empty, skipping..." % patch["name"])
            T.out_best, T.elapsed = T.img * T.mask, 0.
            T.save_result()
            continue
        holders.append(T)
    u.set_seed()                       # the same initial weights on every rank (broadcast from rank 0 anyway)
    outch = a.imgchannel if a.imgchannel is not None else int(patches[0]["image"].shape[-1])
    net = get_net(a, outch).to(dev)
    if a.netdir is not None and len(a.netdir) != 0:
        net.load_state_dict(torch.load(os.path.join("./results", a.netdir[0]), map_location=dev))
    else:
        u.init_weights(net, a.inittype, a.initgain)
    dims = tuple(patches[0]["image"].shape[:-1])
    eng = net.engine_for(dims, dev, max_iters=a.epochs)
    eng.set_loss(a.loss)
    rows = []
    for T in holders:
        T.build_input()
        row = eng.new_row(seed=int(getattr(a, "noise_seed", 0)) * 1000003 + T.patch_index)
        eng.row_load(row, T.input_, T.img_, T.mask_)
        torch.cuda.synchronize(dev)
        T.input_ = None                # the row owns its copy of z now
        rows.append(row)
    sigma = float(a.reg_noise_std) if a.reg_noise_std > 0 else 0.0
    tr = D.SharedNetTrainer(eng, rows, len(live), lr=a.lr, sigma=sigma)
    tr.reset()
    tr.capture()
    opt = FusedAdam(net, lr=a.lr)
    scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=a.lr_factor, threshold=a.lr_thresh,
                                                           patience=a.lr_patience)
    stopper = u.EarlyStopping(patience=a.earlystop_patience, min_delta=a.earlystop_min_delta, percentage=True)
    sync_every = max(1, int(getattr(a, "sync_every", 1)))
    if a.reduce_lr or a.earlystop_patience < a.epochs:
        sync_every = 1
    print("starting optimization with ADAM (shared network, %d rows on this rank, %d in all)..." % (len(rows), len(live)))
    torch.cuda.synchronize(dev)
    start = time()
    j, stop = 0, False
    while j < a.epochs and not stop:
        n = min(sync_every, a.epochs - j)
        for _ in range(n):
            tr.iteration()
        hist = [r.history[j:j + n].cpu().numpy() for r in rows]       # the host <-> device sync of the chunk
        # mean loss over ALL rows: the scalar every rank bases its learning-rate / stopping decisions on
        tot = torch.zeros(n, dtype=torch.float64, device=dev)
        for h in hist:
            tot += torch.from_numpy(h[:, 0].copy()).to(dev)
        D.allreduce_sum_(tot)
        gl = (tot / max(len(live), 1)).cpu().numpy()
        for r in range(n):
            lr_now = opt.param_groups[0]["lr"]
            for T, h in zip(holders, hist):
                T.history.append((float(h[r, 0]), float(h[r, 1]), float(h[r, 2])))
                T.history.lr.append(lr_now)
            if rank == 0:
                print("Iter %s, mean loss over %d rows = %+.2e" % (str(j + r + 1).zfill(u.ten_digit(a.epochs)), len(live),
                                                                    gl[r]), "\r", end="")
            if a.reduce_lr:
                scheduler.step(float(gl[r]))
                if opt.param_groups[0]["lr"] != lr_now:
                    eng.set_lr(opt.param_groups[0]["lr"])
            if stopper.step(float(gl[r])):
                stop = True
                break
        j += n
    torch.cuda.synchronize(dev)
    elapsed = time() - start
    print("\n" + u.sec2time(elapsed))
    for T, row in zip(holders, rows):
        T.net = None
        T.elapsed = elapsed
        T.out_best = T._np_out(eng.row_output_nchw(row))
        T.save_result()
    if rank == 0 and a.savemodel:
        torch.save({k: v.detach().clone() for k, v in net.state_dict().items()}, os.path.join(outpath, "shared_model.pth"))
    if world > 1:
        cs = torch.tensor(tr.param_checksum(), dtype=torch.float64, device=dev)
        both = [torch.zeros_like(cs) for _ in range(world)]
        torch.distributed.all_gather(both, cs)
        if any(not torch.equal(b, both[0]) for b in both):
            raise RuntimeError("shared-network mode: the parameters diverged across ranks")
        torch.distributed.barrier()


def main(argv=None) -> None:
    """``main()`` of main.py:254-297; under torchrun, patches are sharded over ranks (no communication)."""
    warnings.filterwarnings("ignore")
    u.set_seed()
    args = parse_arguments(argv)
    rank, world = _rank_world()
    u.set_gpu(args.gpu)
    if world > 1 and args.outdir is None:
        # every rank would draw its own random result directory (utils.random_code) and reconstruct_patches could not
        # gather the patches again
        raise ValueError("--outdir is required when running under torchrun (WORLD_SIZE > 1)")
    outpath = os.path.join("./results/", args.outdir if args.outdir is not None else u.random_code())
    os.makedirs(outpath, exist_ok=True)
    print("Saving to %s" % outpath)
    if rank == 0:
        u.write_args(os.path.join(outpath, "args.txt"), args)
    patches = extract_patches(args)
    print("Processing %d patches" % len(patches))
    if getattr(args, "shared_net", False):
        _run_shared_net(args, outpath, patches)
        print("Interpolation done! Saved to %s" % outpath)
        return
    if args.start_from_prev and world > 1:
        raise NotImplementedError("--start_from_prev chains patches sequentially (main.py:286); run it on one GPU")
    mine = [(i, patch) for i, patch in enumerate(patches) if i % world == rank]
    k = patches_in_flight(args, patches[0]["image"].shape[:-1], len(mine)) if mine else 1
    if k > 1:
        print("%d patches in flight" % k)
        _run_patches_in_flight(args, outpath, mine, k)
    else:
        T = Interpolator(args, outpath)
        for i, patch in mine:
            if _setup_patch(T, i, patch, args):
                T.optimize()
            T.save_result()
            T.clean()
    print("Interpolation done! Saved to %s" % outpath)


if __name__ == "__main__":
    main()
