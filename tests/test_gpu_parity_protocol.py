"""north_star's parity protocol, literally, on the TF32 tcgen05 path (the path bench.py quotes its number on):

    "per-iteration loss within 1e-3 relative over the first 200 iterations (fp32 accumulate, TF32 operands)"

A free-running deep-prior trajectory is chaotic, so the criterion is checked TEACHER-FORCED at EVERY
one of the first 200 iterations of ``main.py:141-193,210-217``: the oracle (``oracle/net_oracle.py``) runs the loop in
float64 — forward, masked L1, backward, Adam — and at every iteration k the CUDA engine is handed the oracle's current
weights W_k and the same perturbed input z + 0.03 eps_k, runs forward / loss / backward through the C ABI, and its loss
and gradient are compared with the float64 values of that iteration.  Asserted: relative loss error <= 1e-3 at every k.

Gradient: the cosine with the float64 gradient (non-gauge parameters) is reported as min / p5 / median.  How close a
TF32 gradient CAN be depends on the network, not on the kernels (the deepest level of a (32,16,16) patch normalises
over 2 voxels per channel), so the yardstick is ``oracle/tf32_emulation.py`` - the oracle on the CPU with TF32-rounded
conv operands - evaluated on the same weights and input at every iteration: the CUDA path has to be as close to the
float64 gradient as that emulation is (median of 1 - cos within 2x, worst case within 3x), and above an absolute floor.
"""
import json
import os

import numpy as np
import pytest
import torch

from test_gpu_network import FULL, SMALL, gauge_bias, setup

pytestmark = pytest.mark.gpu

ITERS = 200


def _flat_from_state(eng, net, sd):
    """the oracle's float64 state_dict as one fp32 vector in the engine's flat-parameter layout"""
    flat = torch.zeros(eng.params.n, dtype=torch.float32)
    for (name, p), off in zip(net.named_parameters(), eng.params.poff):
        flat[off:off + p.numel()] = sd[name].reshape(-1).float()
    return flat


@pytest.mark.parametrize("widths,dims,cos_floor", [
    (SMALL, (32, 16, 16), 0.99),        # 2 voxels per channel at the deepest level: ill-conditioned BatchNorm
    (FULL, (32, 32, 16), 0.998),        # default widths (the benchmarked network): min 0.99848 here AND in the emulation
])
def test_teacher_forced_200_iterations_tf32(widths, dims, cos_floor):
    from oracle import net_oracle as O
    from oracle import tf32_emulation as T32
    net, sd, z, _, img, mask, cfg = setup("3d", widths, "trilinear", dims, precision="tf32")
    dev = torch.device("cuda")
    sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    z64, img64, mask64 = z.double(), img.double(), mask.double()
    net = net.to(dev)
    eng = net.engine_for(dims, dev, max_iters=8)
    eng.set_noise_input(z.to(dev))
    eng.set_target(img.to(dev), mask.to(dev))
    eng.reset_loop_state(1e-3, 0)
    names = [k for k, _ in net.named_parameters()]
    keep = [k for k in names if not gauge_bias(k, True)]
    sizes = {k: p.numel() for k, p in net.named_parameters()}
    offs = dict(zip(names, eng.params.poff))
    st = O.AdamState()
    g = torch.Generator().manual_seed(2024)
    loss_err, cosines, cos_emu, loss_emu = [], [], [], []

    def cosine(get, g64):
        num = da = db = 0.0
        for name in keep:
            a, b = get(name), g64[name].reshape(-1)
            num += float(a @ b)
            da += float(a @ a)
            db += float(b @ b)
        return num / ((da * db) ** 0.5 + 1e-300)

    for k in range(ITERS):
        e = torch.randn(z.shape, generator=g)
        # ---- CUDA path on the oracle's weights of this iteration ----
        eng.params.P.copy_(_flat_from_state(eng, net, sd64).to(dev))
        eng.perturb_input(0.03, e.to(dev))
        eng.run_forward()
        eng.run_loss()
        eng.run_backward()
        torch.cuda.synchronize()
        l_gpu, _, _ = eng.read_scalars()
        G = eng.params.G.cpu().double()
        # ---- float64 oracle: the same iteration, then its own Adam step (main.py:213) ----
        zin64 = z64 + 0.03 * e.double()
        l_emu, g_emu = T32.loss_and_grads_tf32(sd64, zin64, img64, mask64, cfg, "mae")     # same weights, same input
        l64, _, _, _, g64 = O.loss_and_grads(sd64, zin64, img64, mask64, cfg, "mae")
        with torch.no_grad():
            O.adam_update(sd64, g64, st, lr=1e-3)
        loss_err.append(abs(l_gpu - l64) / abs(l64))
        loss_emu.append(abs(l_emu - l64) / abs(l64))
        cosines.append(cosine(lambda n: G[offs[n]:offs[n] + sizes[n]], g64))
        cos_emu.append(cosine(lambda n: g_emu[n].reshape(-1).double(), g64))
    le, cs, ce, lm = np.array(loss_err), np.array(cosines), np.array(cos_emu), np.array(loss_emu)
    rec = {"iterations": ITERS, "dims": list(dims), "filters": widths["filters"],
           "loss_rel_err": {"max": le.max(), "p95": np.percentile(le, 95), "median": np.median(le)},
           "loss_rel_err_cpu_tf32_emulation": {"max": lm.max(), "p95": np.percentile(lm, 95), "median": np.median(lm)},
           "grad_cosine_k_ge_2": {"min": cs[2:].min(), "p5": np.percentile(cs[2:], 5), "median": np.median(cs[2:])},
           "grad_cosine_cpu_tf32_emulation_k_ge_2": {"min": ce[2:].min(), "p5": np.percentile(ce[2:], 5),
                                                      "median": np.median(ce[2:])},
           "grad_cosine_first_two": [cs[0], cs[1]], "grad_cosine_emulation_first_two": [ce[0], ce[1]]}
    rec = json.loads(json.dumps(rec, default=float))
    print("teacher-forced TF32, 200 iterations:", json.dumps(rec))
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_200it_%s.json" % "x".join(str(f) for f in widths["filters"])), "w") as f:
            json.dump(rec, f, indent=1)
    assert le.max() <= 1e-3, ("north_star: per-iteration loss within 1e-3 relative", int(le.argmax()), le.max())
    d_gpu, d_emu = 1.0 - cs[2:], 1.0 - ce[2:]
    assert np.median(d_gpu) <= 2.0 * np.median(d_emu) + 1e-6, ("median 1-cos: CUDA vs CPU TF32 emulation", np.median(d_gpu), np.median(d_emu))
    assert d_gpu.max() <= 3.0 * d_emu.max() + 1e-6, ("worst 1-cos: CUDA vs CPU TF32 emulation", d_gpu.max(), d_emu.max())
    assert cs[2:].min() >= cos_floor, ("gradient cosine vs the float64 oracle", int(cs[2:].argmin()) + 2, cs[2:].min())
