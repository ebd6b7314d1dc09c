"""CPU oracle of the deep-prior hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product path (``deep_prior_interpolation_b200``) never does.

What it is: a functional restatement, on plain ``torch`` CPU ops, of the one path the reference
(polimi-ispl/deep_prior_interpolation) spends its time in — the MultiRes U-Net forward/backward, the
masked loss, the SNR/PCORR metrics and the Adam update of ``main.py:141-220`` — driven by a flat
``state_dict`` whose keys are the reference's own.  The arithmetic of the
reference lives in a third-party dependency, PyTorch (``environment.yml:13`` pins ``pytorch>=1.7``;
this container has torch 2.11.0), so the restatement calls the same ATen CPU operators the reference's
``nn.Module`` objects dispatch to, but with none of the reference's module code.

Pinning: ``oracle/gen_golden.py`` runs the UNMODIFIED reference from ``/root/reference`` (this container
only) and stores its outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement against those vectors (bit-exact on CPU), and ``tests/test_oracle_vs_reference.py`` checks it
against the live reference when ``/root/reference`` exists.  Parity is therefore pinned on outputs of the
reference itself (the reference ships no tests of its own — SURVEY.md §4).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


@dataclass
class NetConfig:
    """Hyper-parameters of ``get_net`` (architectures/__init__.py:10-86) for the multiunet branch."""
    datadim: str = "3d"                       # '3d' -> MulResUnet3D, '2d'/'2.5d' -> MulResUnet
    inputdepth: int = 64
    outchannel: int = 1
    filters: Sequence[int] = (16, 32, 64, 128, 256)
    skip: Sequence[int] = (16, 32, 64, 128)
    upsample: str = "trilinear"               # 'nearest' | 'bilinear' | 'trilinear'
    activation: str = "LeakyReLU"
    last_activation: Optional[str] = None
    alpha: float = 1.67
    net: str = "multiunet"                    # 'multiunet' | 'attmultiunet' (2-D only, architectures/__init__.py:21-31)

    @property
    def is3d(self) -> bool:
        return self.datadim == "3d"


def block_widths(U: int, alpha: float = 1.67) -> Tuple[int, int, int]:
    """Branch widths of a MultiRes block (mulresunet.py:70-79)."""
    W = alpha * U
    return int(W * 0.167), int(W * 0.333), int(W * 0.5)


def _act(x: torch.Tensor, name: Optional[str]) -> torch.Tensor:
    # architectures/base.py:97-114
    if name is None or name == "none":
        return x
    if name == "LeakyReLU":
        return F.leaky_relu(x, 0.2)
    if name == "ReLU":
        return F.relu(x)
    if name == "ELU":
        return F.elu(x)
    if name == "Tanh":
        return torch.tanh(x)
    if name == "Sigmoid":
        return torch.sigmoid(x)
    raise NotImplementedError(name)


class _Net:
    """Walks the reference's key space; every method cites the module it restates."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg: NetConfig, training: bool = True):
        self.sd, self.cfg, self.training = sd, cfg, training
        self.conv = F.conv3d if cfg.is3d else F.conv2d

    # -- leaves ------------------------------------------------------------------------------------
    def _conv(self, x, key, stride=1):
        # base.py:117-126 / 169-180: same zero padding (k-1)//2
        w = self.sd[key + ".weight"]
        b = self.sd.get(key + ".bias")
        pad = (w.shape[-1] - 1) // 2
        return self.conv(x, w, b, stride=stride, padding=pad)

    def _bn(self, x, key):
        # nn.BatchNorm{2,3}d in training mode: batch statistics, running stats updated in place
        rm, rv = self.sd[key + ".running_mean"], self.sd[key + ".running_var"]
        out = F.batch_norm(x, rm, rv, self.sd[key + ".weight"], self.sd[key + ".bias"],
                           training=self.training, momentum=0.1, eps=1e-5)
        if self.training and (key + ".num_batches_tracked") in self.sd:
            self.sd[key + ".num_batches_tracked"] += 1
        return out

    def _unit(self, x, key):
        # conv3dbn (base.py:211-216): keys <key>.0.0 conv, <key>.1 BN;  conv2dbn (base.py:162-166):
        # <key>.0 conv, <key>.2 BN
        if self.cfg.is3d:
            y = self._bn(self._conv(x, key + ".0.0"), key + ".1")
        else:
            y = self._bn(self._conv(x, key + ".0"), key + ".2")
        return _act(y, self.cfg.activation)

    # -- composite blocks ----------------------------------------------------------------------------
    def block(self, x, key):
        # Block3d.forward (mulresunet.py:85-96) / Block2d.forward (mulresunet.py:27-36)
        o1 = self._unit(x, key + ".conv3x3")
        o2 = self._unit(o1, key + ".conv5x5")
        o3 = self._unit(o2, key + ".conv7x7")
        out = torch.cat([o1, o2, o3], dim=1)
        if self.cfg.is3d:
            out = self._bn(out, key + ".bn1")
        out = torch.add(self._unit(x, key + ".shortcut"), out)
        out = _act(out, self.cfg.activation)
        if self.cfg.is3d:
            out = self._bn(out, key + ".bn2")
        return out

    def respath(self, x, key):
        # ResPath3d.forward (mulresunet.py:108-113) / ResPath2d.forward with length 1 (mulresunet.py:59-64)
        if self.cfg.is3d:
            out = torch.add(self._unit(x, key + ".conv1x1"), self._unit(x, key + ".conv3x3"))
            return self._bn(_act(out, self.cfg.activation), key + ".bn")
        out = torch.add(self._unit(x, key + ".net.0"), self._unit(x, key + ".net.1"))
        return self._bn(_act(out, self.cfg.activation), key + ".net.2")

    def upsample(self, x):
        # nn.Upsample(scale_factor=2, mode=...) (mulresunet.py:168,242)
        mode = self.cfg.upsample
        if mode == "nearest":
            return F.interpolate(x, scale_factor=2, mode="nearest")
        return F.interpolate(x, scale_factor=2, mode=mode)

    @staticmethod
    def crop_cat(a, b):
        # Concat / Concat3D.forward (base.py:296-322, 333-362): centre-crop to the smallest size
        nsp = a.dim() - 2
        tgt = [min(a.shape[2 + i], b.shape[2 + i]) for i in range(nsp)]

        def crop(t):
            sl = [slice(None), slice(None)]
            for i in range(nsp):
                d = (t.shape[2 + i] - tgt[i]) // 2
                sl.append(slice(d, d + tgt[i]))
            return t[tuple(sl)]

        if all(a.shape[2 + i] == b.shape[2 + i] for i in range(nsp)):
            return torch.cat([a, b], dim=1)
        return torch.cat([crop(a), crop(b)], dim=1)

    def level(self, x, key, dec_key, i):
        # one U-Net scale: Concat(skip, deeper) followed by the decoder block (mulresunet.py:216-248)
        n_scales = len(self.cfg.filters)
        if self.cfg.is3d:
            down_bn, enc, main, = key + ".1.2", key + ".1.5", key + ".1.6"
        else:
            down_bn, enc, main, = None, key + ".1.4", key + ".1.5"
        d = self._conv(x, key + ".1.1.0", stride=2)
        if down_bn is not None:
            d = self._bn(d, down_bn)
        d = _act(d, self.cfg.activation)
        e = self.block(d, enc)
        if i < n_scales - 1:
            e = self.level(e, main + ".1", main + ".2", i + 1)
        u = self.upsample(e)
        if self.cfg.skip[i - 1] == 0:
            # mulresunet.py:235-236 drops the Concat (different key space); not part of the hot path
            raise NotImplementedError("skip == 0 is outside the restated path")
        s = self.respath(x, key + ".0.1")
        return self.block(self.crop_cat(s, u), dec_key)

    def forward(self, z):
        x0 = self.block(z, "1")
        y0 = self.level(x0, "2", "3", 1)
        out = self._conv(y0, "4.0")
        la = self.cfg.last_activation
        if isinstance(la, str) and la.lower() == "none":
            la = None
        return _act(out, la)


class _AttNet(_Net):
    """AttMulResUnet2D (architectures/attention.py:197-262) over its own key space: ``down_mb<i>`` / ``up_mb<i>``
    are Block2d, ``down<i>`` = conv(3, stride 2) + BN + act, ``att<i>`` = GridAttentionBlock."""

    def grid_attention(self, g, x, key):
        # GridAttentionBlock.forward (attention.py:107-113): W_g = conv1x1 + BN, W_x = conv3x3 stride 2 + BN,
        # psi = conv1x1 -> Sigmoid -> Upsample(x2, bilinear)
        g1 = self._bn(self._conv(g, key + ".W_g.0.0"), key + ".W_g.1")
        x1 = self._bn(self._conv(x, key + ".W_x.0.0", stride=2), key + ".W_x.1")
        psi = F.relu(g1 + x1)
        psi = torch.sigmoid(self._conv(psi, key + ".psi.0.0"))
        psi = F.interpolate(psi, scale_factor=2, mode="bilinear")
        return x * psi

    def down(self, x, key):
        # attention.py:231-236
        return _act(self._bn(self._conv(x, key + ".0.0", stride=2), key + ".1"), self.cfg.activation)

    def forward(self, z):
        # AttMulResUnet2D.forward (attention.py:249-262)
        n = len(self.cfg.filters)
        xs = [self.block(z, "down_mb1")]
        for i in range(1, n):
            xs.append(self.block(self.down(xs[-1], "down%d" % i), "down_mb%d" % (i + 1)))
        g = xs[-1]
        for i in range(1, n):
            x = xs[n - 1 - i]
            g = self.block(self.crop_cat(self.grid_attention(g, x, "att%d" % i), self.upsample(g)), "up_mb%d" % i)
        la = self.cfg.last_activation
        if isinstance(la, str) and la.lower() == "none":
            la = None
        if la is not None:
            return _act(self._conv(g, "outconv.0.0"), la)
        return self._conv(g, "outconv.0")


def forward(sd: Dict[str, torch.Tensor], z: torch.Tensor, cfg: NetConfig, training: bool = True) -> torch.Tensor:
    """net(input_) of main.py:158 for the multiunet architectures (and the 2-D attention variant)."""
    if cfg.net == "attmultiunet" and not cfg.is3d:
        return _AttNet(sd, cfg, training).forward(z)
    return _Net(sd, cfg, training).forward(z)


def build_state_dict(cfg: NetConfig, seed: int = 0, init_gain: float = 0.02) -> Dict[str, torch.Tensor]:
    """A freshly initialised ``state_dict`` of ``MulResUnet3D`` over the reference's key space,
    built from the constructor arithmetic alone: ``Block3d`` widths (mulresunet.py:67-81), the per-level layout of
    ``MulResUnet3D`` (mulresunet.py:188-259) and the ``init_weights('xavier', 0.02)`` distributions
    (utils/torch.py:34-53: conv ~ xavier_normal(gain), conv bias 0, BatchNorm weight ~ N(10, 10*gain), bias 0).
    Keys / shapes / dtypes are pinned on the reference's own inventory (tests/golden/state_dict_inventory.json); the
    VALUES are equally distributed but not the same draws as ``get_net`` + ``init_weights`` (used where any valid
    initialisation will do: timing the CPU path in bench.py without importing the product package)."""
    if not cfg.is3d or cfg.net != "multiunet":
        raise NotImplementedError("build_state_dict covers MulResUnet3D")
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(key, cin, cout, k):
        fan_in, fan_out = cin * k ** 3, cout * k ** 3
        std = init_gain * math.sqrt(2.0 / (fan_in + fan_out))
        sd[key + ".weight"] = torch.randn((cout, cin, k, k, k), generator=gen) * std
        sd[key + ".bias"] = torch.zeros(cout)

    def bn(key, c):
        sd[key + ".weight"] = 10.0 + 10 * init_gain * torch.randn(c, generator=gen)
        sd[key + ".bias"] = torch.zeros(c)
        sd[key + ".running_mean"] = torch.zeros(c)
        sd[key + ".running_var"] = torch.ones(c)
        sd[key + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)

    def unit(key, cin, cout, k):
        conv(key + ".0.0", cin, cout, k)
        bn(key + ".1", cout)

    def block(key, U, f_in):
        c1, c2, c3 = block_widths(U, cfg.alpha)
        out = c1 + c2 + c3
        unit(key + ".shortcut", f_in, out, 1)
        unit(key + ".conv3x3", f_in, c1, 3)
        unit(key + ".conv5x5", c1, c2, 3)
        unit(key + ".conv7x7", c2, c3, 3)
        bn(key + ".bn1", out)
        bn(key + ".bn2", out)
        return out

    depth = block("1", cfg.filters[0], cfg.inputdepth)
    prefix, dec = "2", "3"
    n = len(cfg.filters)
    for i in range(1, n):
        f = cfg.skip[i - 1]
        unit(prefix + ".0.1.conv3x3", depth, f, 3)
        unit(prefix + ".0.1.conv1x1", depth, f, 1)
        bn(prefix + ".0.1.bn", f)
        conv(prefix + ".1.1.0", depth, depth, 3)
        bn(prefix + ".1.2", depth)
        enc_out = block(prefix + ".1.5", cfg.filters[i], depth)
        block(dec, cfg.filters[i - 1], enc_out + f)
        depth = enc_out
        dec = prefix + ".1.6.2"
        prefix = prefix + ".1.6.1"
    c1, c2, c3 = block_widths(cfg.filters[0], cfg.alpha)
    conv("4.0", c1 + c2 + c3, cfg.outchannel, 3)
    return sd


def param_keys(sd: Dict[str, torch.Tensor]) -> List[str]:
    """state_dict keys that are nn.Parameters (everything except BN buffers), in state_dict order."""
    return [k for k in sd if not (k.endswith("running_mean") or k.endswith("running_var")
                                  or k.endswith("num_batches_tracked"))]


def masked_loss(out, img, mask, kind: str = "mae"):
    """loss_fn(out_ * mask_, img_ * mask_) with reduction='mean' (main.py:24-27,161)."""
    a, b = out * mask, img * mask
    return F.mse_loss(a, b) if kind == "mse" else F.l1_loss(a, b)


def snr(output, target):
    """utils/metrics.py:6-17."""
    return 10 * torch.log10(torch.sum(target ** 2) / torch.sum((target - output) ** 2))


def pcorr(output, target):
    """utils/metrics.py:20-44."""
    mt, mo = torch.mean(target), torch.mean(output)
    td, od = target - mt, output - mo
    return torch.sum(td * od) / (torch.sqrt(torch.sum(td ** 2)) * torch.sqrt(torch.sum(od ** 2)))


@dataclass
class AdamState:
    step: int = 0
    m: Dict[str, torch.Tensor] = field(default_factory=dict)
    v: Dict[str, torch.Tensor] = field(default_factory=dict)


def adam_update(sd, grads: Dict[str, torch.Tensor], st: AdamState, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam.step, single-tensor form (main.py:200,213)."""
    st.step += 1
    bc1, bc2 = 1 - b1 ** st.step, 1 - b2 ** st.step
    for k, g in grads.items():
        if k not in st.m:
            st.m[k] = torch.zeros_like(sd[k])
            st.v[k] = torch.zeros_like(sd[k])
        st.m[k].lerp_(g, 1 - b1)
        st.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (st.v[k].sqrt() / math.sqrt(bc2)).add_(eps)
        sd[k].addcdiv_(st.m[k], denom, value=-(lr / bc1))


def loss_and_grads(sd, z, img, mask, cfg: NetConfig, loss: str = "mae"):
    """One forward + backward of the loop body (main.py:158-167) on detached leaves.

    Returns (loss, snr, pcorr, out, grads-by-key).  BN running statistics in ``sd`` are updated."""
    keys = param_keys(sd)
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in keys}
    work = dict(sd)
    work.update(leaves)
    out = forward(work, z, cfg, training=True)
    l = masked_loss(out, img, mask, loss)
    l.backward()
    grads = {k: (leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])) for k in keys}
    with torch.no_grad():
        s, p = snr(out, img), pcorr(out, img)
    return float(l.detach()), float(s), float(p), out.detach(), grads


def optimisation_iteration(sd, z, noise, img, mask, cfg: NetConfig, st: AdamState, reg_noise_std=0.03, loss="mae",
                           lr=1e-3):
    """Interpolator.optimization_loop + optimizer.step (main.py:141-193, 213) with the per-iteration input
    noise supplied by the caller (``noise`` ~ N(0,1), same shape as z)."""
    zin = z + reg_noise_std * noise if reg_noise_std > 0 else z
    l, s, p, out, grads = loss_and_grads(sd, zin, img, mask, cfg, loss)
    with torch.no_grad():
        adam_update(sd, grads, st, lr=lr)
    return l, s, p, out


def flops_per_voxel_3d() -> float:
    """Training FLOPs per output voxel per iteration of the default MulResUnet3D."""
    return 391285.5
