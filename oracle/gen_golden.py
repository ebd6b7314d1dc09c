"""Generate tests/golden/* by running the UNMODIFIED reference (/root/reference) in this container.

TEST INFRASTRUCTURE.  Run:  python oracle/gen_golden.py
The reference is imported through oracle/refshim.py (stubs for optional third-party imports only); every number
stored here is produced by the reference's own modules: ``architectures.get_net``, ``utils.init_weights``,
``torch.nn.L1Loss/MSELoss``, ``torch.optim.Adam``, ``utils.snr/pcorr``, ``utils.PatchExtractor``, ``utils.bool2bin``,
``parameter.parse_arguments`` — driven exactly like ``main.py:141-213`` drives them.
"""
import json
import os
import sys
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def ref_loop(arch, u, args, dims, loss_kind, iters, seed, outch=1, store_tensors=True):
    """the loop body of main.py:141-213 with the reference's own objects; per-iteration noise from a Generator"""
    torch.manual_seed(seed)
    net = arch.get_net(args, outch)
    u.init_weights(net, "xavier", 0.02)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    z = torch.randn((1, args.inputdepth) + dims, generator=g) * 0.1
    eps = [torch.randn((1, args.inputdepth) + dims, generator=g) for _ in range(iters)]
    img = torch.randn((1, outch) + dims, generator=g) * 2
    tr = (torch.rand((1, 1, 1) + dims[1:], generator=g) > 0.6).float()
    mask = tr.expand((1, outch) + dims).contiguous()
    loss_fn = torch.nn.MSELoss() if loss_kind == "mse" else torch.nn.L1Loss()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    rows, out0, grads0 = [], None, None
    for it in range(iters):
        opt.zero_grad()
        inp = z.detach().clone()
        inp += 0.03 * eps[it]
        out = net(inp)
        total = loss_fn(out * mask, img * mask)
        total.backward()
        rows.append([total.item(), u.snr(output=out, target=img).item(), u.pcorr(output=out, target=img).item()])
        if it == 0:
            out0 = out.detach().clone()
            grads0 = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
        opt.step()
    res = {"rows": np.array(rows, dtype=np.float64), "out0": out0.numpy()}
    res["grad0_norms"] = np.array([float(grads0[k].double().norm()) for k in grads0], dtype=np.float64)
    res["grad0_keys"] = np.array(list(grads0.keys()))
    final = net.state_dict()
    res["final_param_sums"] = np.array([float(final[k].double().sum()) for k in final], dtype=np.float64)
    res["final_param_abs"] = np.array([float(final[k].double().abs().sum()) for k in final], dtype=np.float64)
    res["init_param_abs"] = np.array([float(sd0[k].double().abs().sum()) for k in sd0], dtype=np.float64)
    res["keys"] = np.array(list(final.keys()))
    if store_tensors:
        for k, v in sd0.items():
            res["sd0/" + k] = v.numpy()
        res["z"], res["img"], res["mask"] = z.numpy(), img.numpy(), mask.numpy()
        res["eps"] = torch.stack(eps).numpy()
        # first-iteration gradient of two representative tensors, in full
        for k in ("1.conv3x3.0.0.weight", "1.conv3x3.0.weight", "4.0.weight", "1.bn1.weight", "3.shortcut.1.weight",
                  "3.shortcut.2.weight", "att1.W_x.0.0.weight", "att4.psi.0.0.weight", "att4.W_g.1.weight",
                  "down_mb1.conv3x3.0.weight", "outconv.0.weight"):
            if k in grads0:
                res["grad0/" + k] = grads0[k].numpy()
    return res


def main():
    os.makedirs(OUT, exist_ok=True)
    arch, u, data, parameter, _ = refshim.reference_modules()
    small = dict(inputdepth=8, filters=[4, 8, 16, 32, 64], skip=[4, 8, 16, 32])
    full = dict(inputdepth=64, filters=[16, 32, 64, 128, 256], skip=[16, 32, 64, 128])

    def A(datadim, up, widths, net="multiunet", last=None):
        return Namespace(datadim=datadim, net=net, upsample=up, activation="LeakyReLU", last_activation=last,
                         dropout=0., **widths)

    # the 2-D attention variant (architectures/attention.py:197-262, --net attmultiunet): sizes divisible by 16
    np.savez_compressed(os.path.join(OUT, "attnet2d_small.npz"),
                        **ref_loop(arch, u, A("2d", "bilinear", small, net="attmultiunet"), (48, 32), "mae", 3, 2))
    np.savez_compressed(os.path.join(OUT, "attnet2d_full_scalars.npz"),
                        **ref_loop(arch, u, A("2d", "nearest", full, net="attmultiunet", last="Tanh"), (64, 48), "mse", 2, 0,
                                   store_tensors=False))
    if "--att-only" in sys.argv:
        return
    np.savez_compressed(os.path.join(OUT, "net3d_small.npz"), **ref_loop(arch, u, A("3d", "trilinear", small), (32, 16, 16), "mae", 3, 0))
    np.savez_compressed(os.path.join(OUT, "net3d_small_nearest_mse.npz"),
                        **ref_loop(arch, u, A("3d", "nearest", small), (24, 20, 18), "mse", 2, 3))
    np.savez_compressed(os.path.join(OUT, "net2d_small.npz"), **ref_loop(arch, u, A("2d", "bilinear", small), (43, 25), "mae", 3, 1))
    # default widths: weights are reproducible from the seed (the product's constructors consume the RNG identically),
    # so only scalars / checksums are stored
    np.savez_compressed(os.path.join(OUT, "net3d_full_scalars.npz"),
                        **ref_loop(arch, u, A("3d", "trilinear", full), (32, 16, 16), "mae", 2, 0, store_tensors=False))
    np.savez_compressed(os.path.join(OUT, "net2d_full_scalars.npz"),
                        **ref_loop(arch, u, A("2d", "bilinear", full), (48, 32), "mae", 2, 0, store_tensors=False))

    # state_dict key / shape inventory of the default networks
    inv = {}
    for name, args in (("3d", A("3d", "trilinear", full)), ("2d", A("2d", "bilinear", full))):
        torch.manual_seed(0)
        net = arch.get_net(args, 1)
        inv[name] = [[k, list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()]
        inv[name + "_num_params"] = int(sum(p.numel() for p in net.parameters()))
    with open(os.path.join(OUT, "state_dict_inventory.json"), "w") as f:
        json.dump(inv, f)

    # patches: PatchExtractor.extract / reconstruct, bool2bin
    rng = np.random.RandomState(0)
    pg = {}
    cases = [((23, 17, 11), (8, 6, 5), (4, 3, 5)), ((16, 12, 8), (8, 6, 4), (8, 6, 4)), ((30, 21), (10, 7), (5, 7)),
             ((20, 9, 9), (20, 9, 9), (20, 9, 9))]
    for ci, (shape, dim, stride) in enumerate(cases):
        vol = rng.randn(*shape)
        pe = u.PatchExtractor(dim=dim, stride=stride)
        pa = pe.extract(vol)
        pert = (pa + 0.01 * rng.randn(*pa.shape)).astype(np.float32)
        rec = pe.reconstruct(pert)
        pg["c%d_vol" % ci], pg["c%d_dim" % ci], pg["c%d_stride" % ci] = vol, np.array(dim), np.array(stride)
        pg["c%d_patches" % ci], pg["c%d_pert" % ci], pg["c%d_rec" % ci] = pa, pert, rec
        pg["c%d_rec_gain" % ci] = rec / 40.0
    nanvol = rng.randn(12, 6, 5)
    nanvol[:, rng.rand(6, 5) > 0.5] = np.nan
    pg["nan_in"], pg["nan_out"] = nanvol, u.bool2bin(nanvol)
    np.savez_compressed(os.path.join(OUT, "patches.npz"), **pg)

    # flags: defaults of the reference parser (parameter.py)
    argv = sys.argv
    sys.argv = ["main.py", "--imgdir", "X", "--netdir", "a/b"]
    ns = parameter.parse_arguments()
    sys.argv = ["main.py", "--imgdir", "X", "--netdir", "a/b", "--datadim", "3d", "--upsample", "linear",
                "--patch_shape", "64", "64", "64", "--epochs", "77", "--param_noise", "--savemodel"]
    ns2 = parameter.parse_arguments()
    sys.argv = argv
    with open(os.path.join(OUT, "parse_arguments.json"), "w") as f:
        json.dump({"defaults": vars(ns), "case3d": vars(ns2)}, f, indent=1)

    # host-side helpers
    h = {}
    es = u.EarlyStopping(patience=3, min_delta=1., percentage=True)
    seq = [1.0, 0.995, 0.98, 0.979, 0.978, 0.9779, 0.5]
    h["early_stop"] = [bool(es.step(torch.tensor(v))) for v in seq]
    h["early_stop_seq"] = seq
    h["ten_digit"] = [[n, u.ten_digit(n)] for n in (1, 9, 10, 99, 100, 1470, 2001, 3000)]
    h["sec2time"] = [[s, u.sec2time(s)] for s in (0., 59.9, 61., 3600., 6739.)]
    hist = u.History(3000)
    hist.append((0.5, 1.25, 0.75))
    hist.lr.append(1e-3)
    h["history_msg"] = hist.log_message(0)
    with open(os.path.join(OUT, "host_helpers.json"), "w") as f:
        json.dump(h, f, indent=1)
    print("golden vectors written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print("  %-36s %8.1f KB" % (fn, os.path.getsize(os.path.join(OUT, fn)) / 1024))


if __name__ == "__main__":
    main()
