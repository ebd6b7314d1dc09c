"""Host-side helpers with the semantics of the reference's ``utils`` package (only what the hot path uses).

Cited against ``utils/torch.py``, ``utils/metrics.py``, ``utils/generic.py``, ``utils/processing.py`` and
``utils/mask.py`` of the reference.  Re-implemented, not copied; behaviour (including RNG consumption of
``init_weights``) is kept so seeded runs line up with the reference.
"""
from __future__ import annotations

import json
import math
import os
import random
import string
from argparse import Namespace
from typing import Optional

import numpy as np
import torch

__all__ = ["init_weights", "get_noise", "set_seed", "set_gpu", "get_gpu_name", "np_to_torch", "torch_to_np",
           "EarlyStopping", "snr", "pcorr", "History", "ten_digit", "sec2time", "time2sec", "read_args",
           "write_args", "random_code", "bool2bin", "build_mask", "add_rand_mask", "nextpow2", "lowpass_butterworth_taps",
           "fir_time"]


# ---- utils/torch.py -------------------------------------------------------------------------------------
def init_weights(net: torch.nn.Module, init_type: str = "normal", init_gain: float = 0.02, verbose: bool = False):
    """``init_weights`` (utils/torch.py:23-58): Conv/Linear weights by ``init_type``, zero bias; BatchNorm weight
    ~ N(10, 10*gain), bias 0.  Walks ``net.apply`` so the RNG draws happen in the reference's order."""
    if init_type == "default":
        return

    def visit(m):
        cls = type(m).__name__
        if hasattr(m, "weight") and ("Conv" in cls or "Linear" in cls):
            w = m.weight.data
            if init_type == "normal":
                torch.nn.init.normal_(w, 0.0, init_gain)
            elif init_type == "xavier":
                torch.nn.init.xavier_normal_(w, gain=init_gain)
            elif init_type == "kaiming":
                torch.nn.init.kaiming_normal_(w, a=0.2, mode="fan_in")
            elif init_type == "orthogonal":
                torch.nn.init.orthogonal_(w, gain=init_gain)
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if getattr(m, "bias", None) is not None:
                torch.nn.init.constant_(m.bias.data, 0.0)
        elif "BatchNorm" in cls:
            torch.nn.init.normal_(m.weight.data, 10.0, init_gain * 10)
            torch.nn.init.constant_(m.bias.data, 0.0)

    net.apply(visit)
    if verbose:
        print("parameters initialized with %s" % init_type)


def get_noise(shape, noise_type: str) -> torch.Tensor:
    """``get_noise`` (utils/torch.py:61-73): CPU tensor filled from the default generator."""
    x = torch.zeros(tuple(shape))
    if noise_type == "u":
        x.uniform_()
    elif noise_type == "n":
        x.normal_()
    elif noise_type == "c":
        x.cauchy_()
    else:
        raise ValueError("Noise type has to be one of [u, n, c]")
    return x


def set_seed(seed: int = 0):
    """``set_seed`` (utils/torch.py:198-205) minus the cuDNN switches (no cuDNN on this path)."""
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def set_gpu(id: Optional[int] = -1):
    """``set_gpu`` (utils/torch.py:165-185).  ``None`` is rejected: this implementation has no CPU path."""
    if id is None:
        raise RuntimeError("--gpu is required: deep_prior_interpolation_b200 has no CPU path")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device visible")
    n = torch.cuda.device_count()
    dev = int(os.environ.get("LOCAL_RANK", 0)) if id == -1 else int(id)
    if dev >= n:
        print("The selected GPU does not exist. Switching to device 0.")
        dev = 0
    torch.cuda.set_device(dev)
    print("GPU selected: %d - %s" % (dev, torch.cuda.get_device_name(dev)))
    return dev


def get_gpu_name(id: Optional[int] = None) -> str:
    """``get_gpu_name`` (utils/torch.py:188-195)."""
    if not torch.cuda.is_available():
        return "CPU"
    dev = torch.cuda.current_device() if id is None else int(id)
    return "%s (%d)" % (torch.cuda.get_device_name(dev), dev)


def np_to_torch(a: np.ndarray, bc_add: bool = True) -> torch.Tensor:
    t = torch.from_numpy(a.copy())
    return t.unsqueeze(0).unsqueeze(0) if bc_add else t


def torch_to_np(t: torch.Tensor, bc_del: bool = True) -> np.ndarray:
    a = t.detach().cpu().numpy()
    return a.squeeze() if bc_del else a


class EarlyStopping:
    """``EarlyStopping`` (utils/torch.py:216-275) on host floats: NaN stops, ``patience`` non-improving steps stop."""

    def __init__(self, patience: int = 10, max: bool = False, min_delta: float = 0, percentage: bool = False):
        self.mode = "max" if max else "min"
        self.min_delta, self.patience, self.percentage = min_delta, patience, percentage
        self.best = None
        self.num_bad_epochs = 0
        self.msg = "\nEarly stopping called, terminating..."

    def is_better(self, a, best) -> bool:
        if self.patience == 0:
            return True
        d = best * self.min_delta / 100 if self.percentage else self.min_delta
        return a < best - d if self.mode == "min" else a > best + d

    def step(self, metrics) -> bool:
        if self.patience == 0:
            return False
        m = float(metrics)
        if self.best is None:
            self.best = m
            return False
        if math.isnan(m):
            print("Metrics is NaN, terminating...")
            return True
        if self.is_better(m, self.best):
            self.num_bad_epochs = 0
            self.best = m
        else:
            self.num_bad_epochs += 1
        if self.num_bad_epochs >= self.patience:
            print(self.msg)
            return True
        return False


# ---- utils/metrics.py ------------------------------------------------------------------------------------
def snr(output, target):
    """Signal-to-noise ratio in dB (utils/metrics.py:6-17)."""
    if target.shape != output.shape:
        raise ValueError("There is something wrong with the dimensions!")
    if isinstance(output, torch.Tensor) and isinstance(target, torch.Tensor):
        return 10 * torch.log10(torch.sum(target ** 2) / torch.sum((target - output) ** 2))
    return 10 * np.log10(np.sum(target ** 2) / np.sum((target - output) ** 2))


def pcorr(output, target):
    """Pearson correlation coefficient (utils/metrics.py:20-44)."""
    if target.shape != output.shape:
        raise ValueError("There is something wrong with the dimensions!")
    if isinstance(output, torch.Tensor) and isinstance(target, torch.Tensor):
        td, od = target - torch.mean(target), output - torch.mean(output)
        return torch.sum(td * od) / (torch.sqrt(torch.sum(td ** 2)) * torch.sqrt(torch.sum(od ** 2)))
    td, od = target - np.mean(target), output - np.mean(output)
    return np.sum(td * od) / (np.sqrt(np.sum(td ** 2)) * np.sqrt(np.sum(od ** 2)))


class History:
    """Per-iteration loss / SNR / PCORR / lr lists (utils/metrics.py:47-85); pickled into ``*_run.npy``."""

    def __init__(self, epochs):
        self.loss, self.snr, self.pcorr, self.lr = [], [], [], []
        self.msg = "Iter %s, Loss = %+.2e, SNR = %+2.2f dB, PCORR = %+.2f %%"
        self.zfill = ten_digit(epochs)

    def __getitem__(self, item):
        return self.loss[item], self.snr[item], self.pcorr[item]

    def __setitem__(self, idx, values):
        self.loss[idx], self.snr[idx], self.pcorr[idx] = values

    def append(self, values):
        l, s, p = values
        self.loss.append(l)
        self.snr.append(s)
        self.pcorr.append(p)

    def __len__(self):
        assert len(self.loss) == len(self.snr) == len(self.pcorr) == len(self.lr)
        return len(self.loss)

    def log_message(self, idx):
        return self.msg % (str(idx + 1).zfill(self.zfill), self.loss[idx], self.snr[idx], self.pcorr[idx] * 100)

    def __repr__(self):
        return "Loss : %s\nSNR  : %s\nPCORR: %s" % (self.loss, self.snr, self.pcorr)

    __str__ = __repr__


# ---- utils/generic.py ------------------------------------------------------------------------------------
def nextpow2(x: int) -> int:
    """``nextpow2`` (utils/generic.py:10-11)"""
    return int(math.ceil(math.log2(abs(x))))


# ---- utils/processing.py:34-79 (input-noise pre-filters, run once per patch) ----------------------------------
def lowpass_butterworth_taps(fc: float, fs: float, ntaps: int, order: int, nfft: int) -> np.ndarray:
    """FIR taps of ``LowPassButterworth`` (utils/processing.py:70-79): least-squares fit (``firls``) of ``ntaps``
    taps to the magnitude response of an ``order``-th order digital Butterworth low-pass sampled at ``nfft`` points."""
    from scipy.signal import butter, firls, freqz
    b, a = butter(order, fc, fs=fs, btype="low", analog=False)
    w_iir, h_iir = freqz(b, a, worN=nfft, fs=fs)
    return firls(ntaps, w_iir, abs(h_iir), fs=fs)


def fir_time(x: torch.Tensor, taps: np.ndarray) -> torch.Tensor:
    """``ConvolveKernel_1d(kernel=taps, ndim=x.ndim-2)(x)`` (utils/processing.py:34-67) for a BC[TXY] CUDA tensor: every
    channel is convolved with ``taps`` along the time axis (the first spatial axis), same size, zeros beyond the ends
    (the reference's grouped ``conv_transpose`` with ``padding = len(taps)//2``).  One ``dpi_fir_axis`` launch."""
    import ctypes as C
    from . import _lib
    taps = np.asarray(taps)
    assert taps.ndim == 1 and taps.size % 2 == 1, "odd 1-D kernel expected"
    if not x.is_cuda:
        raise RuntimeError("fir_time needs a CUDA tensor; there is no CPU path")
    x = x.contiguous().float()
    t_dev = torch.from_numpy(taps.astype(np.float32)).to(x.device)      # the reference casts the taps to fp32 too
    y = torch.empty_like(x)
    outer = int(x.shape[0] * x.shape[1])
    T = int(x.shape[2])
    inner = int(np.prod(x.shape[3:])) if x.ndim > 3 else 1
    _lib.call("dpi_fir_axis", C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), outer, T, inner,
              C.c_void_p(t_dev.data_ptr()), int(taps.size), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    return y


def ten_digit(number: float) -> int:
    return int(math.floor(math.log10(number)) + 1)


def sec2time(seconds: float) -> str:
    return "%dh:%dm:%ds" % (seconds // 3600, (seconds // 60) % 60, seconds % 60)


def time2sec(timestamp: str) -> int:
    h, m, s = timestamp.split(":")
    return int(h.replace("h", "")) * 3600 + int(m.replace("m", "")) * 60 + int(s.replace("s", ""))


def random_code(n: int = 6) -> str:
    return "".join(random.choice(string.ascii_letters + string.digits) for _ in range(int(n)))


def read_args(filename) -> Namespace:
    """``read_args`` (utils/generic.py:39-43): args.txt is the JSON dump of the Namespace."""
    args = Namespace()
    with open(filename, "r") as fp:
        args.__dict__.update(json.load(fp))
    return args


def write_args(filename, args: Namespace, indent: int = 2) -> None:
    with open(filename, "w") as fp:
        json.dump(args.__dict__, fp, indent=indent)


# ---- utils/processing.py / utils/mask.py -------------------------------------------------------------------
def bool2bin(in_content: np.ndarray, logic: bool = True) -> np.ndarray:
    """NaN traces -> 0/1 mask (utils/processing.py:27-31)."""
    nan = np.isnan(in_content)
    out = in_content.copy()
    out[~nan] = 1 if logic else 0
    out[nan] = 0 if logic else 1
    return out


def build_mask(data: np.ndarray, rate: float, regular: bool = False) -> np.ndarray:
    """Binary trace-decimation mask (utils/mask.py:6-53); uses ``np.random.choice`` like the reference."""
    if data.ndim == 2:
        nt, nx = data.shape
        ny = 1
    elif data.ndim == 3:
        nt, nx, ny = data.shape
    else:
        raise ValueError("data volume has to be either 2D or 3D")
    ntr = nx * ny
    ndel = int(ntr * rate)
    flat = np.ones((nt, ntr), dtype=data.dtype)
    if regular:
        if rate >= .5:
            keep = ntr - ndel
            m = int(np.ceil(ntr / keep))
            for i in range(keep):
                flat[:, i * m + 1:i * m + m] = 0
        else:
            flat[:] = 0
            m = int(np.ceil(ntr / ndel))
            for i in range(ndel):
                flat[:, i * m + 1:i * m + m] = 1
    else:
        idx = np.random.choice(np.arange(ntr), ndel, replace=False)
        flat[:, idx] = 0
    return flat.reshape((nt, nx, ny)).squeeze()


def add_rand_mask(mask: np.ndarray, perc: float = 0.3) -> np.ndarray:
    """Additive random trace deletion (utils/mask.py:56-75), used by ``--adirandel`` (data.py:79-80)."""
    m = mask.copy()
    live = np.argwhere(m[0] == 1)                      # traces still alive in the first time sample
    pick = np.random.choice(np.arange(live.shape[0]), int(live.shape[0] * perc), replace=False)
    for p in live[pick]:
        if m.ndim == 2:
            m[:, p[0]] = 0
        else:
            m[:, p[0], p[1]] = 0
    return m
