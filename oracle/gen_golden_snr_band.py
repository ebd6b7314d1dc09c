"""Generate tests/golden/snr_band_64.json: the run-to-run band of the UNMODIFIED reference's final SNR.

TEST INFRASTRUCTURE.  Run (this container only, ~1 h of CPU):  python oracle/gen_golden_snr_band.py [iters] [T X Y]

north_star asks for a final reconstruction SNR within 0.2 dB of the reference.  A deep-prior trajectory is chaotic
(the per-iteration input noise alone moves the end point by more than that), so the yardstick is
the reference's OWN spread: the same problem — one synthetic hyperbolic-event patch (``bench.synthetic_patch``, seed
7, 66 % of the traces removed), the same initial weights (``torch.manual_seed(0)``; ``architectures.get_net`` +
``utils.init_weights``) and the same fixed input ``z`` — is optimised by the reference's own modules, driven exactly
like ``main.py:141-213`` drives them (``nn.L1Loss`` on masked tensors, ``torch.optim.Adam(lr=1e-3)``,
``reg_noise_std = 0.03``, ``utils.snr``, best-output tracking of ``main.py:173-182``), once per noise stream.
``tests/test_gpu_snr_band.py`` runs the CUDA path on the same problem and requires its mean final SNR to lie inside
the band recorded here.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (synthetic_patch / default_args only; no product import)
from oracle import refshim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "snr_band_64.json")
NOISE_SEEDS = (11, 12, 13, 14, 15, 16)
# a second file for the converged regime: python oracle/gen_golden_snr_band.py 3000 64 64 64 3  -> snr_band_64_3000.json



def reference_run(arch, u, dims, img, mask, iters, noise_seed):
    args = bench.default_args("fp32")
    torch.manual_seed(0)
    net = arch.get_net(args, 1)
    u.init_weights(net, "xavier", 0.02)
    z = torch.randn((1, 64) + dims) * 0.1
    g = torch.Generator().manual_seed(noise_seed)
    loss_fn = torch.nn.L1Loss()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    L, S = [], []
    best, best_l = None, None
    for it in range(iters):
        opt.zero_grad()
        inp = z.detach().clone()
        inp += 0.03 * torch.randn(z.shape, generator=g)
        out = net(inp)
        total = loss_fn(out * mask, img * mask)
        total.backward()
        l = total.item()
        L.append(l)
        S.append(u.snr(output=out, target=img).item())
        if best_l is None or l <= best_l:
            best_l, best = l, out.detach().clone()
        opt.step()
    return {"noise_seed": noise_seed, "loss": L, "snr": S, "snr_last": S[-1],
            "snr_best_output": float(u.snr(output=best, target=img).item()),
            "snr_mean_last20": float(np.mean(S[-20:])), "loss_last": L[-1]}


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    dims = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (64, 64, 64)
    seeds = NOISE_SEEDS[:int(sys.argv[5])] if len(sys.argv) > 5 else NOISE_SEEDS
    out_path = OUT if iters == 300 else OUT.replace(".json", "_%d.json" % iters)
    arch, u, _, _, _ = refshim.reference_modules()
    img_np, mask_np = bench.synthetic_patch(dims, seed=7)
    img = torch.from_numpy(img_np[..., 0]).float()[None, None]
    mask = torch.from_numpy(mask_np[..., 0]).float()[None, None]
    res = {"what": "final SNR of the unmodified reference (CPU, fp32) on bench.synthetic_patch(dims, seed=7); one run per "
                   "per-iteration noise stream, same initial weights (seed 0) and z",
           "iters": iters, "dims": list(dims), "torch": torch.__version__, "runs": []}
    for s in seeds:
        t0 = time.time()
        r = reference_run(arch, u, dims, img, mask, iters, s)
        r["seconds"] = time.time() - t0
        res["runs"].append(r)
        print("noise stream %d: snr_last %.3f dB, best-output %.3f dB, mean(last 20) %.3f dB, loss %.4e (%.0f s)"
              % (s, r["snr_last"], r["snr_best_output"], r["snr_mean_last20"], r["loss_last"], r["seconds"]), flush=True)
        if iters > 1000:                      # keep the file small: every 10th value of the curves
            r["loss"], r["snr"] = r["loss"][::10], r["snr"][::10]
        with open(out_path, "w") as f:
            json.dump(res, f)
    print("written", out_path)


if __name__ == "__main__":
    main()
