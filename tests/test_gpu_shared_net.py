"""Shared-network mode (BASELINE.json configs[4]; SURVEY.md §8e — not in the reference, whose loop builds one network
per patch, main.py:274-295): ONE network over all patches, batch rows sharded over ranks, one gradient all-reduce per
iteration, identical fused Adam everywhere.

Oracle emulation (local-BN, per row): for every row a forward/backward of the reference arithmetic
(``oracle/net_oracle.py``) on that row alone, gradients averaged over all rows, one Adam step.

* one process, two rows      — the CUDA path against that emulation, teacher-forced over three iterations;
* two processes (two GPUs)   — one row per rank over NCCL: parameters bit-identical across ranks and bit-identical to
                               the one-process two-row run of the same problem (skipped on a one-GPU box).
"""
import os

import numpy as np
import pytest
import torch

from test_gpu_network import SMALL, gauge_bias, make_args

pytestmark = pytest.mark.gpu


def load_run(path):
    from deep_prior_interpolation_b200.data import load_run as _load
    return _load(path)

DIMS = (32, 16, 16)
N_ROWS = 2


def _problem(precision):
    """one network, N_ROWS rows (z, img, mask each) - everything from seeds, so every process builds the same"""
    import deep_prior_interpolation_b200 as dpi
    from deep_prior_interpolation_b200 import utils as u
    from oracle import net_oracle as O
    args = make_args("3d", SMALL, "trilinear", precision)
    torch.manual_seed(0)
    net = dpi.get_net(args, 1)
    u.init_weights(net, "xavier", 0.02)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    rows = []
    for _ in range(N_ROWS):
        z = torch.randn((1, SMALL["inputdepth"]) + DIMS, generator=g) * 0.1
        img = torch.randn((1, 1) + DIMS, generator=g) * 2
        mask = (torch.rand((1, 1, 1) + DIMS[1:], generator=g) > 0.6).float().expand((1, 1) + DIMS).contiguous()
        rows.append((z, img, mask))
    cfg = O.NetConfig(datadim="3d", inputdepth=SMALL["inputdepth"], filters=SMALL["filters"], skip=SMALL["skip"],
                      upsample="trilinear")
    return net, sd, rows, cfg


def _trainer(net, rows, row_ids, n_global, dev, sigma=0.0, max_iters=8):
    from deep_prior_interpolation_b200.distributed import SharedNetTrainer
    net = net.to(dev)
    eng = net.engine_for(DIMS, dev, max_iters=max_iters)
    eng.set_loss("mae")
    rr = []
    for i in row_ids:
        r = eng.new_row(seed=i)
        eng.row_load(r, *(t.to(dev) for t in rows[i]))
        rr.append(r)
    tr = SharedNetTrainer(eng, rr, n_global, lr=1e-3, sigma=sigma)
    tr.reset()
    return net, eng, tr


def test_two_rows_one_process_vs_oracle_emulation():
    from oracle import net_oracle as O
    net, sd, rows, cfg = _problem("fp32")
    dev = torch.device("cuda")
    net, eng, tr = _trainer(net, rows, range(N_ROWS), N_ROWS, dev)
    names = [k for k, _ in net.named_parameters()]
    keep = [k for k in names if not gauge_bias(k, True)]
    offs = dict(zip(names, eng.params.poff))
    sizes = {k: p.numel() for k, p in net.named_parameters()}
    st = O.AdamState()
    for it in range(3):
        # ---- emulation: per-row forward/backward on the same weights, gradients averaged, one Adam step ----
        per_row = [O.loss_and_grads(dict(sd), z, img, mask, cfg, "mae") for (z, img, mask) in rows]
        g_mean = {k: sum(pr[4][k] for pr in per_row) / N_ROWS for k in per_row[0][4]}
        # ---- CUDA path: stage A (rows), no ranks to reduce over, stage B (Adam + bookkeeping) ----
        flat = torch.zeros(eng.params.n)
        for k in names:
            flat[offs[k]:offs[k] + sizes[k]] = sd[k].reshape(-1)
        eng.params.P.copy_(flat.to(dev))
        tr._stage_a()
        torch.cuda.synchronize()
        acc = tr.acc.cpu().double()
        num = da = db = 0.0
        for k in keep:
            a, b = acc[offs[k]:offs[k] + sizes[k]], g_mean[k].double().reshape(-1)
            num += float(a @ b); da += float(a @ a); db += float(b @ b)
        cos = num / (da * db) ** 0.5
        assert cos >= 0.99999, ("mean gradient over the rows vs the emulation", it, cos)
        assert abs(da ** 0.5 / db ** 0.5 - 1) < 1e-4, "gradient scale (1/N pre-scaling)"
        tr._stage_b()
        torch.cuda.synchronize()
        for r, pr in zip(tr.rows, per_row):
            l = float(r.history[it, 0])
            assert abs(l - pr[0]) <= 1e-5 * abs(pr[0]), ("row loss", it, l, pr[0])
        # the same Adam step in the emulation, fed with the CUDA gradient (teacher-forced like test_gpu_network)
        grads = {k: tr.acc[offs[k]:offs[k] + sizes[k]].view(sd[k].shape).cpu().clone() for k in names}
        O.adam_update(sd, grads, st, lr=1e-3)
        P = eng.params.P.cpu()
        for k in names:
            d = (P[offs[k]:offs[k] + sizes[k]] - sd[k].reshape(-1)).abs().max().item()
            assert d <= 2e-7 * (1 + sd[k].abs().max().item()), (it, k, d)
    assert int(eng.counter[0]) == 3 and int(tr.rows[1].counter[0]) == 3
    assert float(eng.hyper[1]) == 4.0, "the Adam step advances once per iteration, not once per row"


def _run_free(rank, world, iters, q=None, port=None):
    """`iters` free-running iterations (graph replay, per-iteration noise on); returns the flat parameters"""
    import torch.distributed as dist
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank if world > 1 else 0)
    net, sd, rows, cfg = _problem("tf32")
    ids = [i for i in range(N_ROWS) if i % world == rank]
    net, eng, tr = _trainer(net, rows, ids, N_ROWS, dev, sigma=0.03, max_iters=iters)
    tr.capture()
    for _ in range(iters):
        tr.iteration()
    torch.cuda.synchronize()
    out = (eng.params.P.cpu().clone(), [r.history[:iters].cpu().clone() for r in tr.rows])
    if world > 1:
        q.put((rank, out[0].numpy(), [h.numpy() for h in out[1]]))
        dist.barrier()
        dist.destroy_process_group()
    return out


def test_two_ranks_equal_one_process_two_rows():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with `gpurun --gpus 2`)")
    import torch.multiprocessing as mp
    iters = 6
    p1, h1 = _run_free(0, 1, iters)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_run_free, args=(r, 2, iters, q, port)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0][1], res[1][1]), "parameters must be bit-identical across ranks"
    assert np.array_equal(res[0][1], p1.numpy()), "two ranks x one row == one process x two rows, bit for bit"
    assert np.array_equal(res[0][2][0], h1[0].numpy()) and np.array_equal(res[1][2][0], h1[1].numpy()), "row histories"
    assert np.isfinite(p1.numpy()).all() and np.isfinite(h1[0].numpy()).all()


def test_shared_net_through_the_command_line(tmp_path, monkeypatch):
    """`--shared_net` through interpolator.main (one process): four 32^3-ish patch rows of one volume share ONE network;
    every row gets its usual <name>_run.npy, reconstruct_patches reassembles them, rank 0 saves shared_model.pth, and
    the mean loss over the rows comes down."""
    from test_gpu_driver import _write_volume
    from deep_prior_interpolation_b200 import interpolator
    from deep_prior_interpolation_b200.data import reconstruct_patches
    monkeypatch.chdir(tmp_path)
    _write_volume(str(tmp_path), (128, 32, 32), 0.5, 7)
    flags = ["--imgdir", str(tmp_path), "--imgname", "original.npy", "--maskname", "decimated.npy", "--datadim", "3d",
             "--gain", "40", "--upsample", "linear", "--patch_shape", "32", "-1", "-1", "--patch_stride", "32", "-1", "-1",
             "--inputdepth", "8", "--filters", "4", "8", "16", "32", "64", "--skip", "4", "8", "16", "32", "--gpu", "0",
             "--epochs", "60", "--savemodel", "--precision", "tf32", "--shared_net", "--outdir", "shared", "--sync_every", "20"]
    interpolator.main(flags)
    out = tmp_path / "results" / "shared"
    files = sorted(os.listdir(out))
    assert files == ["0_run.npy", "1_run.npy", "2_run.npy", "3_run.npy", "args.txt", "shared_model.pth"], files
    runs = [load_run(out / ("%d_run.npy" % p)) for p in range(4)]
    assert len(runs[3]["history"].loss) == 0 and np.abs(runs[3]["output"]).max() < 1e-12      # the empty patch is no row
    for r in runs[:3]:
        assert len(r["history"].loss) == 60 and np.isfinite(r["history"].loss).all()
        assert r["output"].shape == (32, 32, 32) and r["output"].dtype == np.float32
    mean_loss = np.mean([r["history"].loss for r in runs[:3]], axis=0)
    assert mean_loss[40:].min() < 0.8 * mean_loss[0], mean_loss[::10]
    rec = reconstruct_patches(interpolator.parse_arguments(flags))
    assert rec.shape == (128, 32, 32) and np.isfinite(rec).all()
    sd = torch.load(out / "shared_model.pth")
    assert "1.conv3x3.0.0.weight" in sd and all(torch.isfinite(v.float()).all() for v in sd.values())
