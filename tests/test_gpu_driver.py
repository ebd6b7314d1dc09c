"""Drop-in surface on the GPU: patch kernels bit-exact vs the NumPy oracle and the reference's golden vectors, and
the whole driver (`main()` -> args.txt / *_run.npy / *_model.pth -> reconstruct_patches, transfer with --netdir)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def load_run(path):
    from deep_prior_interpolation_b200.data import load_run as _load
    return _load(path)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_patch_kernels_bit_exact_vs_reference_vectors():
    from deep_prior_interpolation_b200.data import PatchExtractor
    g = np.load(os.path.join(GOLD, "patches.npz"))
    for ci in range(4):
        vol, dim, stride = g["c%d_vol" % ci], tuple(int(v) for v in g["c%d_dim" % ci]), tuple(int(v) for v in g["c%d_stride" % ci])
        pe = PatchExtractor(dim=dim, stride=stride)
        pa = pe.extract(vol)
        assert pa.dtype == np.float64 and np.array_equal(pa, g["c%d_patches" % ci]), ci
        rec = pe.reconstruct(g["c%d_pert" % ci])
        assert rec.dtype == np.float32 and np.array_equal(rec, g["c%d_rec" % ci]), ci            # bit-exact
        assert np.array_equal(pe.reconstruct(g["c%d_pert" % ci], gain=40.0), g["c%d_rec_gain" % ci]), ci
        assert np.array_equal(pe.extract(vol, gain=40.0), g["c%d_patches" % ci] * 40.0)


@pytest.mark.parametrize("shape,dim,stride", [((70, 33, 20), (32, 16, 16), (16, 8, 16)), ((64, 64), (64, 64), (64, 64)),
                                              ((100, 40, 24), (64, 32, 24), (32, 8, 24)), ((9, 7, 5), (1, 1, 1), (1, 1, 1))])
def test_patch_roundtrip_properties(shape, dim, stride):
    """extract -> reconstruct of untouched patches returns the cropped volume exactly (averaging equal values);
    compared against the NumPy oracle on ragged shapes, stride == dim (blocks) and 1-voxel patches"""
    from deep_prior_interpolation_b200.data import PatchExtractor
    from oracle import patch_oracle as PO
    rng = np.random.RandomState(1)
    vol = rng.randn(*shape)
    pe = PatchExtractor(dim=dim, stride=stride)
    pa = pe.extract(vol)
    assert np.array_equal(pa, PO.extract(vol, dim, stride))
    f = pa.astype(np.float32)
    rec = pe.reconstruct(f)
    assert np.array_equal(rec, PO.reconstruct(f, dim, stride))
    crop = tuple(slice(0, s) for s in pe.in_content_cropped_shape)
    assert np.allclose(rec, vol[crop].astype(np.float32), rtol=0, atol=1e-6)


def test_patch_errors():
    from deep_prior_interpolation_b200.data import PatchExtractor
    with pytest.raises(ValueError):
        PatchExtractor((8, 8)).extract(np.zeros((4, 4)))            # patch larger than the volume
    with pytest.raises(ValueError):
        PatchExtractor((2, 2, 2)).extract(np.zeros((4, 4)))         # rank mismatch
    pe = PatchExtractor((2, 2), (2, 2))
    pe.extract(np.zeros((4, 4)))
    with pytest.raises(ValueError):
        pe.reconstruct(np.zeros((3, 2, 2, 2), np.float32))          # wrong patch grid


def _write_volume(tmp, shape, rate, seed):
    from deep_prior_interpolation_b200 import utils as u
    rng = np.random.RandomState(seed)
    t = np.arange(shape[0])[:, None, None]
    x = np.arange(shape[1])[None, :, None]
    y = np.arange(shape[2])[None, None, :]
    vol = np.zeros(shape)
    for _ in range(3):
        t0, x0, y0 = rng.uniform(5, shape[0] * 0.7), rng.uniform(0, shape[1]), rng.uniform(0, shape[2])
        tt = np.sqrt(t0 ** 2 + ((x - x0) ** 2 + (y - y0) ** 2) * 0.5)
        a = (np.pi * 0.12 * (t - tt)) ** 2
        vol += 0.1 * (1 - 2 * a) * np.exp(-a)
    np.random.seed(seed)
    mask = u.build_mask(vol, rate)
    dec = vol.copy()
    dec[mask == 0] = np.nan                                         # NaN-trace convention -> bool2bin (data.py:53-54)
    np.save(os.path.join(tmp, "original.npy"), vol)
    np.save(os.path.join(tmp, "decimated.npy"), dec)
    return vol, mask


def test_main_end_to_end_and_transfer(tmp_path, monkeypatch):
    """python -m deep_prior_interpolation_b200.interpolator ... on a synthetic volume, two patches; then a second run
    warm-started from the first run's checkpoints (--netdir, main.py:105-110)"""
    from deep_prior_interpolation_b200 import interpolator, data as D, utils as u
    from deep_prior_interpolation_b200.parameter import parse_arguments
    monkeypatch.chdir(tmp_path)
    vol, mask = _write_volume(str(tmp_path), (64, 16, 16), 0.5, 3)
    common = ["--imgdir", str(tmp_path), "--imgname", "original.npy", "--maskname", "decimated.npy", "--datadim", "3d",
              "--gain", "40", "--upsample", "linear", "--patch_shape", "32", "-1", "-1", "--patch_stride", "32", "-1", "-1",
              "--inputdepth", "8", "--filters", "4", "8", "16", "32", "64", "--skip", "4", "8", "16", "32", "--gpu", "0"]
    interpolator.main(common + ["--outdir", "run1", "--epochs", "40", "--savemodel", "--precision", "tf32"])
    out1 = tmp_path / "results" / "run1"
    assert sorted(os.listdir(out1)) == ["0_model.pth", "0_run.npy", "1_model.pth", "1_run.npy", "args.txt"]
    saved = json.load(open(out1 / "args.txt"))
    assert saved["epochs"] == 40 and saved["upsample"] == "trilinear" and saved["precision"] == "tf32"
    run = load_run(out1 / "0_run.npy")
    assert set(run) == {"device", "elapsed", "outpath", "history", "mask", "image", "output", "noise"}
    h = run["history"]
    assert type(h).__module__ == "utils.metrics" and len(h.loss) == len(h.snr) == len(h.pcorr) == len(h.lr) == 40
    assert run["output"].shape == (32, 16, 16) and run["output"].dtype == np.float32
    assert np.isfinite(h.loss).all() and min(h.loss[20:]) < h.loss[0], "loss should come down within 40 iterations"
    # best output = output of the iteration with the smallest loss (main.py:173-182)
    sd = torch.load(out1 / "0_model.pth")
    assert len(sd) == 448 and sd["4.0.weight"].shape == (1, 6, 3, 3, 3)
    a1 = parse_arguments(common + ["--outdir", "run1"])
    rec = D.reconstruct_patches(a1)
    assert rec.shape == (64, 16, 16) and rec.dtype == np.float32
    assert np.isfinite(rec).all() and np.abs(rec).max() > 0
    # reassembly = per-patch best outputs / gain (patches do not overlap here)
    assert np.array_equal(rec[:32], run["output"] / np.float32(40.0))

    # transfer: second run starts from the saved networks
    interpolator.main(common + ["--outdir", "run2", "--epochs", "5", "--net", "load", "--netdir", "run1/0_model.pth",
                                "run1/1_model.pth"])
    run2 = load_run(tmp_path / "results" / "run2" / "0_run.npy")
    assert run2["history"].loss[0] < h.loss[0], "warm start must begin below the cold start's first loss"


def test_all_zero_patch_is_skipped_and_reassembled(tmp_path, monkeypatch):
    """a patch without any signal is written out as img * mask without optimising (main.py:281-284) - with the channel
    axis the optimised outputs do not have; reconstruct_patches must still reassemble the volume (the reference's
    np.asarray of that ragged list fails).  Found by the 1470-patch run of BASELINE config 4: 300 silent patches."""
    from deep_prior_interpolation_b200 import interpolator, data as D
    from deep_prior_interpolation_b200.parameter import parse_arguments
    monkeypatch.chdir(tmp_path)
    for first_silent in (True, False):
        vol, mask = _write_volume(str(tmp_path), (64, 16, 16), 0.5, 3)
        vol[:32] = 0.0 if first_silent else vol[:32]
        vol[32:] = vol[32:] if first_silent else 0.0
        dec = vol.copy()
        dec[mask == 0] = np.nan
        np.save(tmp_path / "original.npy", vol)
        np.save(tmp_path / "decimated.npy", dec)
        out = "z%d" % int(first_silent)
        argv = ["--imgdir", str(tmp_path), "--imgname", "original.npy", "--maskname", "decimated.npy", "--datadim", "3d",
                "--gain", "40", "--upsample", "linear", "--patch_shape", "32", "-1", "-1", "--patch_stride", "32", "-1", "-1",
                "--inputdepth", "8", "--filters", "4", "8", "16", "32", "64", "--skip", "4", "8", "16", "32", "--gpu", "0",
                "--outdir", out, "--epochs", "5", "--precision", "tf32"]
        interpolator.main(argv)
        silent, live = ("0", "1") if first_silent else ("1", "0")
        rs, rl = load_run(tmp_path / "results" / out / (silent + "_run.npy")), load_run(tmp_path / "results" / out / (live + "_run.npy"))
        assert len(rs["history"].loss) == 0 and np.abs(rs["output"]).max() == 0.0 and rs["output"].shape == (32, 16, 16, 1)
        assert len(rl["history"].loss) == 5 and rl["output"].shape == (32, 16, 16)
        rec = D.reconstruct_patches(parse_arguments(argv))
        assert rec.shape == (64, 16, 16)
        lo, hi = (slice(0, 32), slice(32, 64)) if first_silent else (slice(32, 64), slice(0, 32))
        assert np.abs(rec[lo]).max() == 0.0 and np.array_equal(rec[hi], rl["output"] / np.float32(40.0))


def test_lines_25d_shape_path(tmp_path, monkeypatch):
    """2.5-D mode with 2-D convolutions on a (170,100,1)-shaped volume like datasets/lines (config 2): odd sizes
    170 -> 85 -> 43 -> 22 -> 11 exercise the Concat crop"""
    from deep_prior_interpolation_b200 import interpolator, data as D
    from deep_prior_interpolation_b200.parameter import parse_arguments
    monkeypatch.chdir(tmp_path)
    rng = np.random.RandomState(0)
    t = np.arange(170)[:, None]
    x = np.arange(100)[None, :]
    vol = np.sin(0.2 * (t - 0.5 * x)) * np.exp(-((t - 85) / 60.0) ** 2)
    mask = np.ones_like(vol)
    mask[:, rng.choice(100, 66, replace=False)] = 0
    np.save(tmp_path / "original.npy", vol[..., None])
    np.save(tmp_path / "random66.npy", mask[..., None])
    argv = ["--imgdir", str(tmp_path), "--imgname", "original.npy", "--maskname", "random66.npy", "--datadim", "2.5d",
            "--slice", "tx", "--imgchannel", "1", "--gain", "1", "--upsample", "linear", "--patch_shape", "-1", "-1", "-1",
            "--outdir", "lines", "--epochs", "12", "--gpu", "0", "--inputdepth", "8", "--filters", "4", "8", "16", "32", "64",
            "--skip", "4", "8", "16", "32"]
    interpolator.main(argv)
    run = load_run(tmp_path / "results" / "lines" / "0_run.npy")
    assert run["output"].shape == (170, 100, 1) and np.isfinite(run["history"].loss).all()
    rec = D.reconstruct_patches(parse_arguments(argv))
    assert rec.shape == (170, 100, 1)


def test_patches_in_flight_results_do_not_depend_on_k(tmp_path, monkeypatch):
    """Four independent patches optimised one at a time and three at a time (own stream + CUDA graph each,
    interpolator._run_patches_in_flight) give bit-identical histories, outputs and checkpoints: the scheduler only
    changes what overlaps, not what is computed (the kernels have no atomics and fixed-order reductions)."""
    from deep_prior_interpolation_b200 import interpolator
    monkeypatch.chdir(tmp_path)
    _write_volume(str(tmp_path), (128, 32, 32), 0.5, 7)
    common = ["--imgdir", str(tmp_path), "--imgname", "original.npy", "--maskname", "decimated.npy", "--datadim", "3d",
              "--gain", "40", "--upsample", "linear", "--patch_shape", "32", "-1", "-1", "--patch_stride", "32", "-1", "-1",
              "--inputdepth", "8", "--filters", "4", "8", "16", "32", "64", "--skip", "4", "8", "16", "32", "--gpu", "0",
              "--epochs", "12", "--savemodel", "--precision", "tf32", "--save_every", "5"]
    assert interpolator.patches_in_flight(interpolator.parse_arguments(common), (32, 32, 32), 4) == 1
    assert interpolator.patches_in_flight(interpolator.parse_arguments(common + ["--patches_in_flight", "0"]),
                                          (32, 32, 32), 4) == 3
    interpolator.main(common + ["--outdir", "k1", "--patches_in_flight", "1"])
    interpolator.main(common + ["--outdir", "k3", "--patches_in_flight", "3"])
    d1, d3 = tmp_path / "results" / "k1", tmp_path / "results" / "k3"
    # patch 3 holds no events: it is written out without optimising (main.py:281-284), also from inside the scheduler
    # (its *_model.pth is whatever network the driver object held last, as in the reference: not compared)
    assert sorted(os.listdir(d1)) == sorted(os.listdir(d3)) and len(os.listdir(d1)) == 1 + 3 * 4 + 2
    r = load_run(d3 / "3_run.npy")
    assert np.abs(r["output"]).max() < 1e-12 and len(r["history"].loss) == 0
    for p in range(3):
        r1 = load_run(d1 / ("%d_run.npy" % p))
        r3 = load_run(d3 / ("%d_run.npy" % p))
        assert r1["history"].loss == r3["history"].loss and len(r1["history"].loss) == 12
        assert r1["history"].snr == r3["history"].snr
        assert np.array_equal(r1["output"], r3["output"])
        for it in ("05", "10"):
            assert np.array_equal(np.load(d1 / ("%d_output%s.npy" % (p, it))), np.load(d3 / ("%d_output%s.npy" % (p, it))))
        s1, s3 = torch.load(d1 / ("%d_model.pth" % p)), torch.load(d3 / ("%d_model.pth" % p))
        assert all(torch.equal(s1[k_], s3[k_]) for k_ in s1)


@pytest.mark.parametrize("case", ["c3d", "c25d"])
def test_input_noise_options_match_reference(case, tmp_path):
    """--filter_noise_with_wavelet, --lowpass_fs/fc/ntaps and --data_forgetting_factor (main.py:66-97,153-155) through
    Interpolator.build_input / the engine, against tests/golden/input_options.npz (the unmodified reference driven by
    oracle/gen_golden_input_options.py): filtered noise, normalised data tensor, the first network input, and the
    losses of the loop with the reference's own per-iteration noise."""
    import input_options_common as X
    from deep_prior_interpolation_b200.interpolator import Interpolator
    from deep_prior_interpolation_b200.parameter import parse_arguments
    c, g = X.CASES[case], X.golden()
    np.save(tmp_path / "wavelet.npy", g["wavelet"])
    flags = ["--imgdir", str(tmp_path), "--datadim", c["datadim"], "--upsample", "linear", "--imgchannel", str(c["outch"]),
             "--loss", c["loss"], "--lowpass_fs", str(c["fs"]), "--lowpass_fc", str(c["fc"]), "--lowpass_ntaps", str(c["ntaps"]),
             "--data_forgetting_factor", str(c["factor"]), "--inputdepth", "8", "--filters", "4", "8", "16", "32", "64",
             "--skip", "4", "8", "16", "32", "--epochs", str(c["iters"]), "--gpu", "0"]
    if c["wavelet"]:
        flags.append("--filter_noise_with_wavelet")
    args = parse_arguments(flags)
    dev = torch.device("cuda")
    img, mask = X.synthetic(c["dims"], c["outch"])
    T = Interpolator(args, str(tmp_path))
    T.load_data({"image": img, "mask": mask, "name": "0"})
    T.net = X.initial_net(c).to(dev)               # the reference's initial weights (same seed, CPU generator)
    torch.manual_seed(X.SEED_Z)
    T.build_input()
    ref = g[case + "/input_filtered"]
    assert np.abs(T.input_.cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()
    ref = g[case + "/add_data"]
    assert np.abs(T.add_data_.cpu().numpy() - ref).max() <= 2e-6 * np.abs(ref).max()
    assert np.allclose(T.add_data_weight, g[case + "/add_data_weight"], rtol=1e-14)

    eng = T.net.engine_for(tuple(T.input_.shape[2:]), dev, max_iters=8)
    eng.set_loss(c["loss"])
    eng.set_noise_input(T.input_)
    eng.set_target(T.img_, T.mask_)
    eng.set_data_forgetting(T.add_data_, T.add_data_weight)
    eng.reset_loop_state(1e-3, 0)
    rows = []
    for it in range(c["iters"]):
        eng.perturb_input(0.03, X.eps_of(it, T.input_.shape).to(dev))
        eng.add_forgetting_data()
        if it == 0:
            ref = g[case + "/net_input0"]
            assert np.abs(eng.network_input_nchw().cpu().numpy()[0] - ref).max() <= 1e-6 * np.abs(ref).max()
        eng.run_forward()
        eng.run_loss()
        eng.run_backward()
        eng.adam_step()
        eng.iteration_end()
        rows.append(eng.read_scalars())
    rows, ref = np.array(rows), g[case + "/rows"]
    # Iteration 0 (same weights, same input) to round-off.  Later iterations run on weights that went through Adam,
    # whose first step moves EVERY parameter by lr * sign(gradient): where the gradient is rounding noise the sign
    # depends on summation order, so free-running trajectories are compared loosely (tests/test_oracle_golden.py; the
    # 1e-3 per-iteration criterion is checked teacher-forced in tests/test_gpu_network.py)
    assert np.allclose(rows[0, 0], ref[0, 0], rtol=2e-5, atol=0), (rows[:, 0], ref[:, 0])
    assert np.allclose(rows[:, 0], ref[:, 0], rtol=2e-2, atol=0), (rows[:, 0], ref[:, 0])
    assert np.allclose(rows[0, 1], ref[0, 1], rtol=0, atol=2e-4)

    # the public path (graph replay, Philox noise): the inputs of the first F iterations are recorded (main.py:155),
    # the data term fades out with logspace(0,-4,F), and after F iterations the input is z + 0.03*eps again
    T.optimize()
    assert len(T.input_list) == c["factor"] and len(T.history.loss) == c["iters"] and np.isfinite(T.history.loss).all()
    z = T.input_.cpu().numpy()[0]
    rep = T.add_data_.cpu().numpy()[0][np.arange(8) % c["outch"]]
    for it in range(c["factor"]):
        resid = T.input_list[it] - z - np.float32(T.add_data_weight[it]) * rep
        assert abs(resid.std() - 0.03) < 2e-3 and abs(resid.mean()) < 2e-3, (it, resid.std(), resid.mean())
    resid = eng.network_input_nchw().cpu().numpy()[0] - z
    assert abs(resid.std() - 0.03) < 2e-3
