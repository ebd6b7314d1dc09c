"""Shared set-up of the input-option parity tests (CPU oracle test and GPU product test): re-creates, from the seeds
used by oracle/gen_golden_input_options.py, the data / weights / noise the reference consumed for
tests/golden/input_options.npz."""
import os
from argparse import Namespace

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED_NET, SEED_Z, SEED_EPS = 5, 6, 1000
SMALL = dict(inputdepth=8, filters=[4, 8, 16, 32, 64], skip=[4, 8, 16, 32])
CASES = {
    "c3d": dict(datadim="3d", upsample="trilinear", dims=(32, 16, 16), outch=1, loss="mae", iters=4, factor=3,
                wavelet=True, fs=250., fc=30., ntaps=7),
    "c25d": dict(datadim="2.5d", upsample="bilinear", dims=(43, 25), outch=3, loss="mse", iters=3, factor=2,
                 wavelet=False, fs=100., fc=20., ntaps=11),
}


def golden():
    return np.load(os.path.join(GOLD, "input_options.npz"), allow_pickle=False)


def synthetic(dims, outch, seed=7):
    rng = np.random.RandomState(seed)
    img = rng.randn(*dims, outch) * 2.0
    tr = (rng.rand(*((1,) + tuple(dims[1:]) + (1,))) > 0.6).astype(np.float64)
    return img, np.broadcast_to(tr, img.shape).copy()


def to_bc(a: np.ndarray) -> torch.Tensor:
    """(t, x[, y], c) float64 -> (1, c, t, x[, y]) float32, as Interpolator.load_data does (main.py:131-135)"""
    axes = tuple(range(a.ndim))
    return torch.from_numpy(np.transpose(a, axes[-1:] + axes[:-1]).copy()).unsqueeze(0).float()


def noise_z(c):
    torch.manual_seed(SEED_Z)
    return torch.zeros((1, SMALL["inputdepth"]) + c["dims"]).normal_() * 0.1


def eps_of(it, shape):
    torch.manual_seed(SEED_EPS + it)
    return torch.zeros(shape).normal_()


def net_args(c, precision="fp32"):
    return Namespace(datadim=c["datadim"], net="multiunet", upsample=c["upsample"], activation="LeakyReLU",
                     last_activation=None, dropout=0., precision=precision, **SMALL)


def initial_net(c, precision="fp32"):
    """same weights as the reference's build_model under torch.manual_seed(SEED_NET) (constructors and init_weights
    consume the CPU generator identically: tests/test_host_cpu.py)"""
    import deep_prior_interpolation_b200 as dpi
    from deep_prior_interpolation_b200 import utils as u
    torch.manual_seed(SEED_NET)
    net = dpi.get_net(net_args(c, precision), c["outch"])
    u.init_weights(net, "xavier", 0.02)
    return net
