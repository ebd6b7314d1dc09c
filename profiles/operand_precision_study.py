#!/usr/bin/env python
"""CPU study for the next round: what would bf16 conv operands / bf16 activation storage cost in accuracy?

Teacher-forced (same weights, same input) loss error and gradient cosine against a float64 evaluation of the oracle,
for three emulated arithmetic modes of the convolutions (accumulation always fp32):
  tf32  - conv operands (activations, weights, output gradients) rounded to 10 mantissa bits (what the tcgen05
          kind::tf32 path of this round does)
  bf16  - conv operands rounded to bf16 (kind::f16 MMAs, K = 16 per instruction: half the MMAs, half the operand bytes)
  bf16s - additionally every conv OUTPUT is stored as bf16 (so BatchNorm statistics / normalisation read bf16): the
          "all activations in bf16" layout that would halve the HBM traffic of the streaming kernels

    python profiles/operand_precision_study.py [--full]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import net_oracle as O  # noqa: E402


def round_mant(x: torch.Tensor, bits: int) -> torch.Tensor:
    """round-to-nearest-even to `bits` explicit mantissa bits (fp32 in, fp32 out)"""
    if bits >= 23:
        return x
    if bits == 7:
        return x.to(torch.bfloat16).to(torch.float32)
    i = x.view(torch.int32)
    drop = 23 - bits
    half = (1 << (drop - 1)) - 1
    lsb = (i >> drop) & 1
    return ((i + half + lsb) >> drop << drop).view(torch.float32)


class _RoundBoth(torch.autograd.Function):
    """value rounded on the way forward, gradient rounded on the way back"""

    @staticmethod
    def forward(ctx, x, fbits, bbits):
        ctx.bbits = bbits
        return round_mant(x, fbits)

    @staticmethod
    def backward(ctx, g):
        return round_mant(g.contiguous(), ctx.bbits), None, None


def patched_net(mode):
    op_bits = {"fp32": 23, "tf32": 10, "bf16": 7, "bf16s": 7}[mode]
    store_bits = 7 if mode == "bf16s" else 23

    class Net(O._Net):
        def _conv(self, x, key, stride=1):
            w = self.sd[key + ".weight"]
            b = self.sd.get(key + ".bias")
            pad = (w.shape[-1] - 1) // 2
            if x.dtype == torch.float64:
                return self.conv(x, w, b, stride=stride, padding=pad)
            xr = _RoundBoth.apply(x, op_bits, 23)
            wr = _RoundBoth.apply(w, op_bits, 23)
            y = self.conv(xr, wr, b, stride=stride, padding=pad)
            # dy is an MMA operand of dgrad / wgrad: rounded on the way back; y itself stored at `store_bits`
            return _RoundBoth.apply(y, store_bits, op_bits)

    return Net


def evaluate(sd, z, img, mask, cfg, loss, mode):
    keys = O.param_keys(sd)
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in keys}
    work = {k: v.clone() for k, v in sd.items()}
    work.update(leaves)
    net = patched_net(mode)(work, cfg, True)
    out = net.forward(z)
    l = O.masked_loss(out, img, mask, loss)
    l.backward()
    g = torch.cat([leaves[k].grad.reshape(-1).double() for k in keys if leaves[k].grad is not None and not k.endswith(".bias")])
    return float(l.detach()), g, out.detach()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="default widths (inputdepth 64, filters 16..256) instead of the small net")
    a = ap.parse_args()
    import deep_prior_interpolation_b200 as dpi
    from deep_prior_interpolation_b200 import utils as u
    from argparse import Namespace
    widths = dict(inputdepth=64, filters=[16, 32, 64, 128, 256], skip=[16, 32, 64, 128]) if a.full else \
        dict(inputdepth=8, filters=[4, 8, 16, 32, 64], skip=[4, 8, 16, 32])
    dims = (32, 32, 32)
    args = Namespace(datadim="3d", net="multiunet", upsample="trilinear", activation="LeakyReLU", last_activation=None,
                     dropout=0., **widths)
    cfg = O.NetConfig(datadim="3d", inputdepth=widths["inputdepth"], outchannel=1, filters=tuple(widths["filters"]),
                      skip=tuple(widths["skip"]), upsample="trilinear", activation="LeakyReLU", last_activation=None)
    print("net: %s, patch %s" % ("default widths" if a.full else "small", dims))
    for trial in range(2):
        torch.manual_seed(trial)
        net = dpi.get_net(args, 1)
        u.init_weights(net, "xavier", 0.02)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        g = torch.Generator().manual_seed(100 + trial)
        z = torch.randn((1, widths["inputdepth"]) + dims, generator=g) * 0.1 + 0.03 * torch.randn((1, widths["inputdepth"]) + dims, generator=g)
        img = torch.randn((1, 1) + dims, generator=g) * 2
        mask = (torch.rand((1, 1, 1) + dims[1:], generator=g) > 0.6).float().expand((1, 1) + dims).contiguous()
        sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
        l64, g64, o64 = evaluate(sd64, z.double(), img.double(), mask.double(), cfg, "mae", "fp32")
        for mode in ("fp32", "tf32", "bf16", "bf16s"):
            l, gg, o = evaluate(sd, z, img, mask, cfg, "mae", mode)
            cos = float((gg @ g64) / (gg.norm() * g64.norm()))
            oe = float((o.double() - o64).abs().max() / o64.abs().max())
            print("trial %d  %-5s loss rel.err %.2e   gradient cosine %.6f   output max rel.err %.2e"
                  % (trial, mode, abs(l - l64) / abs(l64), cos, oe))


if __name__ == "__main__":
    main()
