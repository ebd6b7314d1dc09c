"""Free-running comparison (north_star: final SNR within 0.2 dB of the reference): the same synthetic hyperbolic
patch is optimised for N iterations by (a) the CPU oracle = reference arithmetic, two different per-iteration noise
streams, (b) this implementation in fp32 and tf32, two noise seeds each.  Free-running trajectories are chaotic
(SURVEY.md fact 12), so the reference's own run-to-run band is printed next to our numbers.
usage: python scratch/snr_compare.py [iters] [T X Y]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import deep_prior_interpolation_b200 as dpi  # noqa: E402
from deep_prior_interpolation_b200 import utils as u  # noqa: E402
from oracle import net_oracle as O  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dims = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (64, 32, 32)
img_np, mask_np = bench.synthetic_patch(dims, seed=7)
img = torch.from_numpy(img_np[..., 0]).float()[None, None]
mask = torch.from_numpy(mask_np[..., 0]).float()[None, None]
args = bench.default_args("fp32")
res = {"iters": iters, "dims": dims, "runs": []}


def final_stats(hist_loss, hist_snr, out_best):
    s_best = float(u.snr(out_best, img))
    return {"loss_last": hist_loss[-1], "snr_last": hist_snr[-1], "snr_best_output": s_best,
            "snr_mean_last20": float(np.mean(hist_snr[-20:]))}


def run_oracle(noise_seed):
    torch.manual_seed(0)
    net = dpi.get_net(args, 1)
    u.init_weights(net, "xavier", 0.02)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    z = torch.randn((1, 64) + dims) * 0.1
    g = torch.Generator().manual_seed(noise_seed)
    st = O.AdamState()
    cfg = O.NetConfig()
    L, S, best, best_l = [], [], None, None
    for it in range(iters):
        l, s, p, out = O.optimisation_iteration(sd, z, torch.randn(z.shape, generator=g), img, mask, cfg, st, 0.03, "mae", 1e-3)
        L.append(l)
        S.append(s)
        if best_l is None or l <= best_l:
            best_l, best = l, out.clone()
    return final_stats(L, S, best)


def run_gpu(precision, noise_seed):
    a = bench.default_args(precision)
    torch.manual_seed(0)
    net = dpi.get_net(a, 1)
    u.init_weights(net, "xavier", 0.02)
    z = torch.randn((1, 64) + dims) * 0.1
    dev = torch.device("cuda")
    net = net.to(dev)
    eng = net.engine_for(dims, dev, max_iters=iters)
    eng.set_noise_input(z.to(dev))
    eng.set_target(img.to(dev), mask.to(dev))
    eng.reset_loop_state(1e-3, noise_seed)
    eng.capture(0.03, 0)
    for _ in range(iters):
        eng.graph.replay()
    torch.cuda.synchronize()
    h = eng.history[:iters].cpu().numpy()
    return final_stats(list(h[:, 0]), list(h[:, 1]), eng.output_nchw(best=True).cpu())


t0 = time.time()
for seed in (11, 12):
    r = run_oracle(seed)
    r["who"] = "reference arithmetic (CPU oracle, fp32), noise stream %d" % seed
    res["runs"].append(r)
    print(json.dumps(r), flush=True)
print("oracle time %.0fs" % (time.time() - t0), flush=True)
for prec in ("fp32", "tf32"):
    for seed in (1, 2):
        r = run_gpu(prec, seed)
        r["who"] = "B200 %s, Philox seed %d" % (prec, seed)
        res["runs"].append(r)
        print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "snr_compare.json"), "w") as f:
    json.dump(res, f, indent=1)
