"""CPU restatement (NumPy) of the reference's input-noise options — TEST INFRASTRUCTURE, never imported by the product.

Follows main.py:66-97 (set-up in ``Interpolator.build_input``) and main.py:153-155 (use inside ``optimization_loop``):
  * ``fir_time``            — ``ConvolveKernel_1d.forward`` (utils/processing.py:34-67)
  * ``butterworth_taps``    — ``LowPassButterworth.__init__`` (utils/processing.py:70-79)
  * ``forgetting_data``     — main.py:86-97
  * ``network_input``       — main.py:147-155
Pinned by tests/test_oracle_golden.py against tests/golden/input_options.npz, which oracle/gen_golden_input_options.py
produced by running the unmodified reference.
"""
import numpy as np


def fir_time(x: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """x: (B, C, T, ...) float32.  The reference builds a kernel that is zero except for ``taps`` along the time axis and
    calls a grouped ``conv_transpose`` with padding = len(taps)//2, i.e. a true convolution of every channel along T,
    cropped to the input length:  y[t] = sum_m taps[m] * x[t + pad - m]."""
    taps = np.asarray(taps, dtype=np.float32)
    n = taps.size
    pad = n // 2
    T = x.shape[2]
    y = np.zeros_like(x, dtype=np.float32)
    for m in range(n):
        s = pad - m                         # y[t] += taps[m] * x[t + s]
        lo, hi = max(0, -s), min(T, T - s)
        if hi > lo:
            y[:, :, lo:hi] += taps[m] * x[:, :, lo + s:hi + s]
    return y


def butterworth_taps(fc: float, fs: float, ntaps: int, nfft: int, order: int = 4) -> np.ndarray:
    from scipy.signal import butter, firls, freqz
    b, a = butter(order, fc, fs=fs, btype="low", analog=False)
    w, h = freqz(b, a, worN=nfft, fs=fs)
    return firls(ntaps, w, abs(h), fs=fs)


def forgetting_data(img: np.ndarray, mask: np.ndarray, inp: np.ndarray, factor: int):
    """img, mask: (1, Cimg, ...) float32; inp: the (filtered) noise tensor (1, inputdepth, ...).  Returns the image
    channels of the normalised data tensor (channel c of the full tensor is channel c % Cimg) and the weights."""
    data = (img * mask).astype(np.float32)
    depth = inp.shape[1]
    rep = int(np.ceil(depth / data.shape[1]))
    full = np.tile(data, (1, rep) + (1,) * (data.ndim - 2))[:, :depth]
    ratio = np.float32(np.std(inp.astype(np.float64), ddof=1) / np.std(full.astype(np.float64), ddof=1))   # torch.std is unbiased
    return data * ratio, np.logspace(0, -4, factor)


def network_input(z: np.ndarray, eps: np.ndarray, sigma: float, add_data: np.ndarray = None, weight: float = 0.0):
    """input_ = z + sigma * eps  [+ weight * repeat(add_data)]  in float32, in the reference's order of operations"""
    out = z.astype(np.float32) + np.float32(sigma) * eps.astype(np.float32)
    if add_data is not None:
        c = np.arange(z.shape[1]) % add_data.shape[1]
        out = out + np.float32(weight) * add_data[:, c]
    return out
