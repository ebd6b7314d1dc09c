"""The oracle with TF32 convolution operands, on the CPU — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The tcgen05 path multiplies TF32 operands (10 explicit mantissa bits; activations, weights and output gradients are
rounded before they reach the tensor core) and accumulates in fp32.  This module evaluates ``oracle/net_oracle.py``
with exactly that operand rounding applied around every convolution (forward: input and weight; backward: the output
gradient, which is the MMA operand of dgrad and wgrad) and everything else in fp32.  It answers "how far from the
float64 truth is ANY TF32 implementation of this network?", which is the yardstick the gradient-cosine floors of the
GPU tests are set against (tests/test_gpu_parity_protocol.py).
"""
import torch

from . import net_oracle as O


def round_mant(x: torch.Tensor, bits: int) -> torch.Tensor:
    """round-to-nearest-even to `bits` explicit mantissa bits (fp32 in, fp32 out)"""
    if bits >= 23:
        return x
    i = x.contiguous().view(torch.int32)
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


class _Tf32Net(O._Net):
    def _conv(self, x, key, stride=1):
        w = self.sd[key + ".weight"]
        b = self.sd.get(key + ".bias")
        pad = (w.shape[-1] - 1) // 2
        xr = _RoundBoth.apply(x, 10, 23)
        wr = _RoundBoth.apply(w, 10, 23)
        y = self.conv(xr, wr, b, stride=stride, padding=pad)
        return _RoundBoth.apply(y, 23, 10)


def loss_and_grads_tf32(sd, z, img, mask, cfg: O.NetConfig, loss: str = "mae"):
    """(loss, grads-by-key) of one forward + backward with TF32 conv operands; ``sd`` is not modified"""
    keys = O.param_keys(sd)
    leaves = {k: sd[k].detach().clone().float().requires_grad_(True) for k in keys}
    work = {k: (v.clone().float() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    work.update(leaves)
    out = _Tf32Net(work, cfg, True).forward(z.float())
    l = O.masked_loss(out, img.float(), mask.float(), loss)
    l.backward()
    grads = {k: (leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])) for k in keys}
    return float(l.detach()), grads
