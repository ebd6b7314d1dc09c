"""Golden vectors for the input-noise options of main.py:66-97,153-155, produced by the UNMODIFIED
reference: ``--filter_noise_with_wavelet``, ``--lowpass_fs/--lowpass_fc/--lowpass_ntaps``, ``--data_forgetting_factor``.

TEST INFRASTRUCTURE.  Run (in the container that has /root/reference):  python oracle/gen_golden_input_options.py

The reference's own ``Interpolator`` (main.py:18-220) is driven on the CPU: ``load_data`` -> ``build_model`` ->
``build_input`` -> ``optimization_loop`` + ``Adam.step`` a few times.  Everything random is drawn from seeded torch CPU
generators in a fixed order, so the test re-creates weights / z / eps from the seeds and only the reference's RESULTS
are stored: the filtered input tensor, the normalised data tensor and its weights, the first perturbed network input,
and the per-iteration loss / SNR / PCORR.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

SEED_NET, SEED_Z, SEED_EPS = 5, 6, 1000


def ricker(n: int, f: float = 0.12) -> np.ndarray:
    t = np.arange(n) - n // 2
    a = (np.pi * f * t) ** 2
    return ((1 - 2 * a) * np.exp(-a)).astype(np.float64)


def synthetic(dims, outch, seed):
    """(image, mask) in the (t, x[, y], channel) layout data.extract_patches hands to Interpolator.load_data"""
    rng = np.random.RandomState(seed)
    img = rng.randn(*dims, outch) * 2.0
    tr = (rng.rand(*((1,) + tuple(dims[1:]) + (1,))) > 0.6).astype(np.float64)
    return img, np.broadcast_to(tr, img.shape).copy()


def drive(main, parameter, flags, dims, outch, iters, tmpdir):
    argv = sys.argv
    sys.argv = ["main.py", "--imgdir", tmpdir] + flags
    args = parameter.parse_arguments()
    sys.argv = argv
    args.gpu = None
    args.netdir = []                  # the reference crashes on len(None) without --netdir (main.py:105)
    args.param_noise = False          # (a no-op on the weights in the reference, but it consumes RNG draws)
    img, mask = synthetic(dims, outch, 7)
    T = main.Interpolator(args, tmpdir)
    T.load_data({"image": img, "mask": mask, "name": "0"})
    torch.manual_seed(SEED_NET)
    T.build_model()
    torch.manual_seed(SEED_Z)
    T.build_input()
    T.optimizer = torch.optim.Adam(T.net.parameters(), lr=args.lr)
    res = {"input_filtered": T.input_.numpy().copy()}
    if T.add_data_ is not None:
        res["add_data"] = T.add_data_[:, :outch].numpy().copy()
        res["add_data_weight"] = np.asarray(T.add_data_weight, dtype=np.float64)
        # every channel block of add_data_ repeats the image channels (main.py:91)
        for c in range(T.add_data_.shape[1]):
            assert torch.equal(T.add_data_[:, c], T.add_data_[:, c % outch])
    for it in range(iters):
        torch.manual_seed(SEED_EPS + it)
        eps = torch.empty_like(T.input_).normal_()        # the draw optimization_loop makes first (main.py:150)
        torch.manual_seed(SEED_EPS + it)
        T.optimizer.zero_grad()
        T.optimization_loop()
        T.optimizer.step()
        if it < args.data_forgetting_factor:
            want = T.input_ + args.reg_noise_std * eps
            want = want + float(np.float32(T.add_data_weight[it])) * T.add_data_
            got = torch.from_numpy(T.input_list[it])[None]
            assert torch.equal(want, got), "eps replay does not reproduce the reference's network input"
            if it == 0:
                res["net_input0"] = T.input_list[0].copy()
    res["rows"] = np.stack([np.array(T.history.loss), np.array(T.history.snr), np.array(T.history.pcorr)], 1).astype(np.float64)
    res["n_inputs_recorded"] = np.array(len(T.input_list))
    return res


def main():
    import tempfile
    _, u, _, parameter, main_mod = refshim.reference_modules()
    tmp = tempfile.mkdtemp()
    np.save(os.path.join(tmp, "wavelet.npy"), ricker(9))
    small = ["--inputdepth", "8", "--filters", "4", "8", "16", "32", "64", "--skip", "4", "8", "16", "32", "--epochs", "4"]
    out = {}
    r = drive(main_mod, parameter, small + ["--datadim", "3d", "--upsample", "linear", "--imgchannel", "1",
                                            "--filter_noise_with_wavelet", "--lowpass_fs", "250", "--lowpass_fc", "30",
                                            "--data_forgetting_factor", "3"], (32, 16, 16), 1, 4, tmp)
    out.update({"c3d/" + k: v for k, v in r.items()})
    r = drive(main_mod, parameter, small + ["--datadim", "2.5d", "--upsample", "linear", "--imgchannel", "3", "--loss", "mse",
                                            "--lowpass_fs", "100", "--lowpass_fc", "20", "--lowpass_ntaps", "11",
                                            "--data_forgetting_factor", "2"], (43, 25), 3, 3, tmp)
    out.update({"c25d/" + k: v for k, v in r.items()})
    out["wavelet"] = ricker(9)
    # the FIR taps the reference derives from the Butterworth response (utils/processing.py:70-79)
    out["taps_3d"] = u.LowPassButterworth(fc=30., ndim=3, fs=250., ntaps=7, order=4, nfft=2 ** u.nextpow2(32),
                                          dtype=torch.FloatTensor).taps
    out["taps_25d"] = u.LowPassButterworth(fc=20., ndim=2, fs=100., ntaps=11, order=4, nfft=2 ** u.nextpow2(43),
                                           dtype=torch.FloatTensor).taps
    path = os.path.join(OUT, "input_options.npz")
    np.savez_compressed(path, **out)
    print("written %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))
    for k, v in out.items():
        print("  %-28s %s" % (k, getattr(v, "shape", v)))
    print("c3d rows:\n", out["c3d/rows"])
    print("c25d rows:\n", out["c25d/rows"])


if __name__ == "__main__":
    main()
