"""CPU tests that PIN THE ORACLE: ``oracle/net_oracle.py`` / ``oracle/patch_oracle.py`` against the golden
vectors produced by the unmodified reference (``oracle/gen_golden.py`` -> ``tests/golden``), and — when
``/root/reference`` is present — against the live reference.  These run without a GPU."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import net_oracle as O
from oracle import patch_oracle as PO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL = dict(inputdepth=8, filters=(4, 8, 16, 32, 64), skip=(4, 8, 16, 32))


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def _run_oracle(g, cfg, loss, iters):
    sd = {k[4:]: torch.from_numpy(g[k].copy()) for k in g.files if k.startswith("sd0/")}
    keys = [str(k) for k in g["keys"]]
    sd = {k: sd[k] for k in keys}                       # reference key order
    z, img, mask = (torch.from_numpy(g[n].copy()) for n in ("z", "img", "mask"))
    eps = torch.from_numpy(g["eps"].copy())
    st = O.AdamState()
    rows = []
    out0 = grads0 = None
    for it in range(iters):
        zin = z + 0.03 * eps[it]
        l, s, p, out, grads = O.loss_and_grads(sd, zin, img, mask, cfg, loss)
        if it == 0:
            out0, grads0 = out, grads
        with torch.no_grad():
            O.adam_update(sd, grads, st, lr=1e-3)
        rows.append([l, s, p])
    return sd, np.array(rows), out0, grads0


@pytest.mark.parametrize("fname,datadim,up,loss,iters", [
    ("net3d_small.npz", "3d", "trilinear", "mae", 3),
    ("net3d_small_nearest_mse.npz", "3d", "nearest", "mse", 2),
    ("net2d_small.npz", "2d", "bilinear", "mae", 3),
    ("attnet2d_small.npz", "2d", "bilinear", "mae", 3),          # --net attmultiunet (attention.py:197-262)
])
def test_oracle_reproduces_reference_loop(fname, datadim, up, loss, iters):
    """losses / SNR / PCORR of every iteration, the first output, first gradients and the parameters after the
    Adam steps — all bit-comparable with what the reference's own modules produced (same ATen CPU kernels)."""
    g = _load(fname)
    cfg = O.NetConfig(datadim=datadim, upsample=up, net="attmultiunet" if fname.startswith("att") else "multiunet",
                      **SMALL)
    sd, rows, out0, grads0 = _run_oracle(g, cfg, loss, iters)
    ref = g["rows"]
    # Iterations 0 and 1 agree to fp32 round-off.  From iteration 2 on the reference's trajectory is driven by
    # rounding NOISE: 48 of 49 conv biases feed a BatchNorm, their true gradient is 0, and Adam turns the noise
    # into +-lr steps whose sign depends on summation order (measured: those biases differ by 2*lr after ONE
    # step between two evaluations of the same graph).  Free-running values are therefore only compared loosely
    # (SURVEY.md fact 12); the per-iteration 1e-3 criterion is checked teacher-forced elsewhere.
    assert np.allclose(rows[:2, 0], ref[:2, 0], rtol=1e-6, atol=0), (rows[:, 0], ref[:, 0])
    assert np.allclose(rows[:, 0], ref[:, 0], rtol=2e-2, atol=0), (rows[:, 0], ref[:, 0])
    assert np.allclose(rows[:2, 1], ref[:2, 1], rtol=0, atol=1e-4)
    assert np.allclose(rows[:2, 2], ref[:2, 2], rtol=0, atol=1e-5)
    assert np.abs(out0.numpy() - g["out0"]).max() <= 1e-6 * np.abs(g["out0"]).max()
    gk = [str(k) for k in g["grad0_keys"]]
    norms = np.array([float(grads0[k].double().norm()) for k in gk])
    big = g["grad0_norms"] > 1e-6 * g["grad0_norms"].max()
    assert np.allclose(norms[big], g["grad0_norms"][big], rtol=1e-4)
    for k in g.files:
        if k.startswith("grad0/"):
            ref_g = g[k]
            assert np.abs(grads0[k[6:]].numpy() - ref_g).max() <= 1e-5 * np.abs(ref_g).max() + 1e-12, k
    keys = [str(k) for k in g["keys"]]
    sums = np.array([float(sd[k].double().abs().sum()) for k in keys])
    # every element moved by at most iters*lr; where the gradient is above the noise both sides moved the same way
    budget = np.array([0.2 * iters * 1e-3 * sd[k].numel() for k in keys])
    sel = np.array([k.endswith(".weight") for k in keys])
    assert np.all(np.abs(sums - g["final_param_abs"])[sel] <= budget[sel] + 1e-5 * g["final_param_abs"][sel]), \
        "weights after the Adam steps"


@pytest.mark.parametrize("fname,datadim,up,dims,kind,last,loss", [
    ("net3d_full_scalars.npz", "3d", "trilinear", (32, 16, 16), "multiunet", None, "mae"),
    ("net2d_full_scalars.npz", "2d", "bilinear", (48, 32), "multiunet", None, "mae"),
    ("attnet2d_full_scalars.npz", "2d", "nearest", (64, 48), "attmultiunet", "Tanh", "mse")])
def test_default_width_network_from_seed(fname, datadim, up, dims, kind, last, loss):
    """default widths: the product's constructors consume the RNG exactly like the reference's, so the same seed
    gives the same initial weights; the oracle then reproduces the reference's two iterations."""
    deep = pytest.importorskip("deep_prior_interpolation_b200")
    from deep_prior_interpolation_b200 import utils as u
    from argparse import Namespace
    g = _load(fname)
    args = Namespace(datadim=datadim, net=kind, upsample=up, activation="LeakyReLU", last_activation=last,
                     dropout=0., inputdepth=64, filters=[16, 32, 64, 128, 256], skip=[16, 32, 64, 128])
    torch.manual_seed(0)
    net = deep.get_net(args, 1)
    u.init_weights(net, "xavier", 0.02)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    keys = [str(k) for k in g["keys"]]
    assert list(sd.keys()) == keys
    init_abs = np.array([float(sd[k].double().abs().sum()) for k in keys])
    assert np.allclose(init_abs, g["init_param_abs"], rtol=1e-12), "initial weights differ from the reference's"
    gen = torch.Generator().manual_seed(1)
    z = torch.randn((1, 64) + dims, generator=gen) * 0.1
    eps = [torch.randn((1, 64) + dims, generator=gen) for _ in range(2)]
    img = torch.randn((1, 1) + dims, generator=gen) * 2
    tr = (torch.rand((1, 1, 1) + dims[1:], generator=gen) > 0.6).float()
    mask = tr.expand((1, 1) + dims).contiguous()
    cfg = O.NetConfig(datadim=datadim, upsample=up, net=kind, last_activation=last)
    st = O.AdamState()
    rows = []
    for it in range(2):
        l, s, p, _ = O.optimisation_iteration(sd, z, eps[it], img, mask, cfg, st, 0.03, loss, 1e-3)
        rows.append([l, s, p])
    rows = np.array(rows)
    assert np.allclose(rows[:, 0], g["rows"][:, 0], rtol=1e-6), (rows, g["rows"])
    assert np.allclose(rows[:, 1:], g["rows"][:, 1:], atol=1e-4)


def test_patch_oracle_against_reference_vectors():
    g = _load("patches.npz")
    for ci in range(4):
        vol, dim, stride = g["c%d_vol" % ci], tuple(g["c%d_dim" % ci]), tuple(g["c%d_stride" % ci])
        pa = PO.extract(vol, dim, stride)
        assert pa.shape == g["c%d_patches" % ci].shape and np.array_equal(pa, g["c%d_patches" % ci])
        rec = PO.reconstruct(g["c%d_pert" % ci], dim, stride)
        assert rec.dtype == np.float32 and np.array_equal(rec, g["c%d_rec" % ci])          # bit-exact
        assert np.array_equal(rec / 40.0, g["c%d_rec_gain" % ci])
    assert np.array_equal(PO.bool2bin(g["nan_in"]), g["nan_out"])


def test_live_reference_if_present():
    """when the reference checkout is available (this container), compare on fresh random inputs too"""
    from oracle import refshim
    if not refshim.available():
        pytest.skip("/root/reference not present (GPU box)")
    arch, u, data, parameter, _ = refshim.reference_modules()
    from argparse import Namespace
    for datadim, up, dims, kind in (("3d", "trilinear", (24, 18, 20), "multiunet"), ("2d", "nearest", (37, 29), "multiunet"),
                                    ("2d", "bilinear", (32, 48), "attmultiunet")):
        args = Namespace(datadim=datadim, net=kind, upsample=up, activation="LeakyReLU", last_activation=None,
                         dropout=0., inputdepth=8, filters=[4, 8, 16, 32, 64], skip=[4, 8, 16, 32])
        torch.manual_seed(5)
        net = arch.get_net(args, 2)
        u.init_weights(net, "xavier", 0.02)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        z = torch.randn((1, 8) + dims)
        out_ref = net(z)
        cfg = O.NetConfig(datadim=datadim, upsample=up, outchannel=2, net=kind, **SMALL)
        out = O.forward(sd, z, cfg)
        assert torch.equal(out, out_ref.detach()), (datadim, kind)
    rng = np.random.RandomState(3)
    vol = rng.randn(19, 14, 9)
    pe = u.PatchExtractor(dim=(6, 5, 4), stride=(3, 4, 2))
    pa = pe.extract(vol)
    assert np.array_equal(pa, PO.extract(vol, (6, 5, 4), (3, 4, 2)))
    f = pa.astype(np.float32)
    assert np.array_equal(pe.reconstruct(f), PO.reconstruct(f, (6, 5, 4), (3, 4, 2)))


# ---- input-noise options (main.py:66-97,153-155) ------------------------------------------------
@pytest.mark.parametrize("case", ["c3d", "c25d"])
def test_input_options_oracle_matches_reference(case):
    """oracle/input_options_oracle.py against what the unmodified reference produced (oracle/gen_golden_input_options.py):
    filtered noise, normalised data tensor, first network input, and the losses of the loop run on those inputs."""
    from oracle import input_options_oracle as IO
    import input_options_common as X
    c, g = X.CASES[case], X.golden()
    img, mask = (X.to_bc(a) for a in X.synthetic(c["dims"], c["outch"]))
    z = X.noise_z(c).numpy()
    if c["wavelet"]:
        z = IO.fir_time(z, g["wavelet"])
    taps = IO.butterworth_taps(c["fc"], c["fs"], c["ntaps"], nfft=2 ** int(np.ceil(np.log2(c["dims"][0]))))
    assert np.allclose(taps, g["taps_3d" if case == "c3d" else "taps_25d"], rtol=1e-12, atol=1e-15)
    z = IO.fir_time(z, taps)
    ref = g[case + "/input_filtered"]
    assert np.abs(z - ref).max() <= 1e-6 * np.abs(ref).max()
    add, w = IO.forgetting_data(img.numpy(), mask.numpy(), z, c["factor"])
    assert np.allclose(w, g[case + "/add_data_weight"], rtol=1e-14)
    assert np.abs(add - g[case + "/add_data"]).max() <= 2e-6 * np.abs(add).max()
    x0 = IO.network_input(z, X.eps_of(0, z.shape).numpy(), 0.03, add, w[0])
    assert np.abs(x0[0] - g[case + "/net_input0"]).max() <= 1e-6 * np.abs(x0).max()
    assert int(g[case + "/n_inputs_recorded"]) == c["factor"]
    # The loop, fed the reference's own filtered noise and data tensor (bit-identical inputs).  Iterations 0 and 1
    # must agree to round-off; from iteration 2 on free-running trajectories are driven by the rounding-noise
    # gradients of the BatchNorm-fed conv biases (see test_oracle_reproduces_reference_loop) and are compared loosely.
    z, add = g[case + "/input_filtered"], g[case + "/add_data"]
    net = X.initial_net(c)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    cfg = O.NetConfig(datadim=c["datadim"], inputdepth=8, outchannel=c["outch"], filters=tuple(X.SMALL["filters"]),
                      skip=tuple(X.SMALL["skip"]), upsample=c["upsample"], activation="LeakyReLU", last_activation=None)
    st = O.AdamState()
    rows = []
    for it in range(c["iters"]):
        e = X.eps_of(it, z.shape).numpy()
        zin = IO.network_input(z, e, 0.03, add if it < c["factor"] else None, w[it] if it < c["factor"] else 0.0)
        l, s, p, out, grads = O.loss_and_grads(sd, torch.from_numpy(zin), img, mask, cfg, c["loss"])
        with torch.no_grad():
            O.adam_update(sd, grads, st, lr=1e-3)
        rows.append([l, s, p])
    rows, ref = np.array(rows), g[case + "/rows"]
    assert np.allclose(rows[:2, 0], ref[:2, 0], rtol=1e-6, atol=0), (rows[:, 0], ref[:, 0])
    assert np.allclose(rows[:, 0], ref[:, 0], rtol=2e-2, atol=0), (rows[:, 0], ref[:, 0])
    assert np.allclose(rows[:2, 1], ref[:2, 1], rtol=0, atol=1e-4)
    assert np.allclose(rows[:2, 2], ref[:2, 2], rtol=0, atol=1e-5)


def test_build_state_dict_matches_reference_inventory():
    """oracle.build_state_dict (used by bench.py's CPU arm instead of the product's constructors): keys, shapes and
    dtypes equal the unmodified reference's MulResUnet3D state_dict (tests/golden/state_dict_inventory.json)"""
    import json
    import os
    from oracle import net_oracle as O
    inv = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "state_dict_inventory.json")))
    sd = O.build_state_dict(O.NetConfig(), seed=0)
    assert {k: (tuple(v.shape), str(v.dtype)) for k, v in sd.items()} == {k: (tuple(sh), dt) for k, sh, dt in inv["3d"]}
    assert sum(sd[k].numel() for k in O.param_keys(sd)) == inv["3d_num_params"]
    w = sd["2.0.1.conv3x3.0.0.weight"]          # xavier_normal(gain 0.02): std = 0.02 * sqrt(2 / (fan_in + fan_out))
    assert abs(float(w.std()) / (0.02 * (2.0 / ((25 + 16) * 27)) ** 0.5) - 1) < 0.05
    assert abs(float(sd["1.bn1.weight"].mean()) - 10.0) < 0.2
