"""Network constructors with the reference's surface: ``get_net``, ``MulResUnet3D``, ``MulResUnet``.

Drop-in contract:

* same constructor signatures as ``architectures/mulresunet.py:116-126,188-198`` and
  ``architectures/__init__.py:10``;
* ``state_dict()`` has exactly the reference's keys / shapes / dtypes, so ``*_model.pth`` files load both
  ways (``main.py:108-110,238-240``);
* the module tree uses stock ``nn.Conv*d`` / ``nn.BatchNorm*d`` objects purely as PARAMETER HOLDERS, created
  in the reference's construction order, so the same ``torch.manual_seed`` yields bit-identical initial
  weights and ``init_weights`` (class-name matching, ``utils/torch.py:34-53``) behaves identically;
* ``net(input_)`` maps ``(1, inputdepth, T, X, Y)`` -> ``(1, outchannel, T, X, Y)`` and is differentiable
  through ``torch.autograd`` (``loss.backward()`` fills ``p.grad``).

None of the holder modules' ``forward`` is ever called: ``forward`` of the top-level container runs the
compiled CUDA plan of ``engine.Engine`` (hand-written sm_100a kernels behind ``include/dpi_b200.h``).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
from torch import nn

__all__ = ["get_net", "MulResUnet", "MulResUnet3D", "AttMulResUnet2D", "DeepPriorNet"]

_ACTS = ("LeakyReLU", "ReLU", "ELU", "Tanh", "Sigmoid")


def _act_module(name: str) -> nn.Module:
    # architectures/base.py:97-114 — placeholders only (keep child numbering identical)
    if name == "LeakyReLU":
        return nn.LeakyReLU(0.2, inplace=True)
    if name == "ELU":
        return nn.ELU()
    if name == "none":
        return nn.Sequential()
    if name == "ReLU":
        return nn.ReLU()
    if name == "Tanh":
        return nn.Tanh()
    if name == "Sigmoid":
        return nn.Sigmoid()
    raise NotImplementedError("unknown activation function %r" % (name,))


def _widths(U: int, alpha: float):
    W = alpha * U
    return int(W * 0.167), int(W * 0.333), int(W * 0.5)


class _Numbered(nn.Sequential):
    """Container whose children are named '1','2',... like the reference's ``Module.add`` (base.py:69-73)."""

    def put(self, m: nn.Module):
        self.add_module(str(len(self) + 1), m)
        return m


class _Holder(nn.Module):
    """Named-children container (Block / ResPath / Concat); never executed."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("holder modules are parameter containers; call the top-level network")


class _Builder:
    def __init__(self, is3d: bool, act: str, bias: bool, dropout: float):
        self.is3d, self.act, self.bias, self.dropout = is3d, act, bias, dropout
        self.Conv = nn.Conv3d if is3d else nn.Conv2d
        self.BN = nn.BatchNorm3d if is3d else nn.BatchNorm2d
        self.Drop = nn.Dropout3d if is3d else nn.Dropout2d

    def conv(self, cin, cout, k, stride=1):
        """returns (container, conv): ``conv()`` / ``conv3d()`` of base.py:117-126,169-180"""
        c = self.Conv(cin, cout, k, stride, padding=(k - 1) // 2, bias=self.bias)
        return nn.Sequential(c), c

    def unit(self, cin, cout, k):
        """conv-BN-act unit; returns (container, (conv, bn)) — conv3dbn / conv2dbn (base.py:162-166,211-216)"""
        seq, c = self.conv(cin, cout, k)
        bn = self.BN(cout)
        if self.is3d:
            box = nn.Sequential(seq, bn, _act_module(self.act))          # keys 0.0 / 1 / 2
        else:
            box = seq                                                      # keys 0 / 2 / 3
            box.add_module("2", bn)
            box.add_module("3", _act_module(self.act))
        return box, (c, bn)

    def block(self, U, cin, alpha):
        """MultiRes block holder (Block3d / Block2d, mulresunet.py:11-36,67-96)"""
        c1, c2, c3 = _widths(U, alpha)
        h = _Holder()
        spec = {}
        for name, (a, b, k) in (("shortcut", (cin, c1 + c2 + c3, 1)), ("conv3x3", (cin, c1, 3)),
                                ("conv5x5", (c1, c2, 3)), ("conv7x7", (c2, c3, 3))):
            box, cb = self.unit(a, b, k)
            h.add_module(name, box)
            spec[name] = cb
        if self.is3d:
            h.add_module("bn1", self.BN(c1 + c2 + c3))
            h.add_module("bn2", self.BN(c1 + c2 + c3))
            spec["bn1"], spec["bn2"] = h.bn1, h.bn2
        else:
            h.add_module("dr", self.Drop(self.dropout))
        h.add_module("act", _act_module(self.act))
        if self.is3d:
            h.add_module("dr", self.Drop(self.dropout))
        h.out_dim = c1 + c2 + c3
        return h, spec

    def respath(self, cin, cout):
        """ResPath3d / ResPath2d(length=1) holder (mulresunet.py:39-64,99-113)"""
        h = _Holder()
        if self.is3d:
            b3, cb3 = self.unit(cin, cout, 3)
            b1, cb1 = self.unit(cin, cout, 1)
            h.add_module("conv3x3", b3)
            h.add_module("conv1x1", b1)
            h.add_module("bn", self.BN(cout))
            h.add_module("act", _act_module(self.act))
            h.add_module("dr", self.Drop(self.dropout))
            return h, {"conv3x3": cb3, "conv1x1": cb1, "bn": h.bn}
        h.add_module("dr", self.Drop(self.dropout))
        b3, cb3 = self.unit(cin, cout, 3)
        b1, cb1 = self.unit(cin, cout, 1)
        bn = self.BN(cout)
        h.add_module("act", _act_module(self.act))
        h.add_module("net", nn.Sequential(b3, b1, bn, h.dr))
        return h, {"conv3x3": cb3, "conv1x1": cb1, "bn": bn}


class DeepPriorNet(_Numbered):
    """Top-level container: reference-identical module tree + the compiled B200 plan as its ``forward``."""

    def _init_runtime(self, spec, precision: str):
        object.__setattr__(self, "spec", spec)
        object.__setattr__(self, "_engine", None)
        object.__setattr__(self, "precision", precision)

    # -- engine management ----------------------------------------------------------------------------
    def engine_for(self, spatial: Sequence[int], device, max_iters: int = 4096):
        from .engine import Engine
        dims = tuple(int(s) for s in spatial)
        dims3 = dims if len(dims) == 3 else (1,) + dims
        eng = self._engine
        if eng is None or eng.dims != dims3 or eng.device != torch.device(device) or eng.max_iters < max_iters:
            if self.spec["dropout"] != 0.0:
                raise NotImplementedError("dropout > 0 is not part of the accelerated path (parameter.py:42 default 0)")
            object.__setattr__(self, "_engine", None)   # free the old plan's buffers first
            eng = Engine(self, dims3, device, precision=self.precision, max_iters=max_iters)
            object.__setattr__(self, "_engine", eng)
        elif eng.params.stale():
            eng.params.adopt()
        return eng

    def release_engine(self):
        object.__setattr__(self, "_engine", None)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("deep_prior_interpolation_b200 runs on CUDA (sm_100a) only; got a CPU tensor — "
                               "there is no CPU fallback")
        want_nd = 5 if self.spec["is3d"] else 4
        if self.spec.get("kind") == "attmultiunet" and x.dim() == 4:
            div = 2 ** len(self.spec["levels"])           # 16 for the default five scales
            if any(int(n) % div for n in x.shape[2:]):
                # the reference fails in `x * psi` (attention.py:113) when a level has an odd size: RuntimeError there too
                raise RuntimeError("attmultiunet needs spatial sizes divisible by %d (got %s): the up-sampled attention "
                                   "map must match the skip tensor (attention.py:109-113)" % (div, tuple(x.shape[2:])))
        if x.dim() != want_nd or x.shape[0] != 1 or x.shape[1] != self.spec["inputdepth"]:
            raise ValueError("expected input of shape (1, %d, %s), got %s" % (
                self.spec["inputdepth"], "T, X, Y" if self.spec["is3d"] else "H, W", tuple(x.shape)))
        eng = self.engine_for(x.shape[2:], x.device)
        params = list(self.parameters())
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _NetFunction.apply(x, self, *params)
        eng.set_network_input(x)
        eng.run_forward()
        return eng.output_nchw()


class _NetFunction(torch.autograd.Function):
    """Bridges the compiled plan into torch.autograd so ``loss.backward()`` (main.py:162) keeps working."""

    @staticmethod
    def forward(ctx, x, net, *params):
        eng = net._engine
        eng.set_network_input(x.detach())
        eng.run_forward()
        ctx.net = net
        return eng.output_nchw()

    @staticmethod
    def backward(ctx, grad_out):
        net = ctx.net
        eng = net._engine
        eng._to_cl(grad_out, eng.out.gptr, eng.out_layout, eng.out.ld)
        eng.run_backward()
        # gradients live in the flat buffer: p.grad becomes a view of it (or is accumulated into when the
        # caller kept an unrelated .grad tensor alive)
        P = eng.params
        for p, off in zip(P.plist, P.poff):
            v = P.G[off:off + p.numel()].view(p.shape)
            if p.grad is None:
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                p.grad.add_(v)
        return (None, None) + (None,) * len(P.plist)


def _build(is3d: bool, num_input_channels, num_output_channels, num_channels_down, num_channels_up,
           num_channels_skip, alpha, last_act_fun, need_bias, upsample_mode, act_fun, dropout, precision):
    assert len(num_channels_down) == len(num_channels_up) == (len(num_channels_skip) + 1)
    if any(s == 0 for s in num_channels_skip):
        raise NotImplementedError("num_channels_skip entries must be non-zero (reference default 16 32 64 128)")
    n_scales = len(num_channels_down)
    if isinstance(upsample_mode, (list, tuple)):
        upsample_mode = upsample_mode[0]
    if act_fun not in _ACTS:
        raise NotImplementedError("unknown activation function %r" % (act_fun,))
    b = _Builder(is3d, act_fun, need_bias, dropout)
    Concat = type("Concat3D" if is3d else "Concat", (_Holder,), {})

    model = DeepPriorNet()
    first, first_spec = b.block(num_channels_down[0], num_input_channels, alpha)
    model.put(first)
    levels = []
    input_depth = first.out_dim
    parent = model
    for i in range(1, n_scales):
        # construction order follows mulresunet.py:216-248 so the RNG stream of the default init matches
        enc, enc_spec = b.block(num_channels_down[i], input_depth, alpha)
        down_box, down_conv = b.conv(input_depth, input_depth, 3, stride=2)
        down_bn = b.BN(input_depth) if is3d else None
        rp, rp_spec = b.respath(input_depth, num_channels_skip[i - 1])
        deeper, skip = _Numbered(), _Numbered()
        deeper.put(down_box)
        if down_bn is not None:
            deeper.put(down_bn)
        deeper.put(_act_module(act_fun))
        deeper.put(b.Drop(dropout))
        deeper.put(enc)
        skip.put(rp)
        cat = Concat()
        cat.dim = 1
        cat.add_module("0", skip)
        cat.add_module("1", deeper)
        parent.put(cat)
        deeper_main = _Numbered()
        if i != n_scales - 1:
            deeper.put(deeper_main)
        deeper.put(nn.Upsample(scale_factor=2, mode=upsample_mode))
        dec, dec_spec = b.block(num_channels_up[i - 1], enc.out_dim + num_channels_skip[i - 1], alpha)
        parent.put(dec)
        levels.append({"down": (down_conv, down_bn), "enc": enc_spec, "respath": rp_spec, "dec": dec_spec})
        input_depth = enc.out_dim
        parent = deeper_main
    out_box, out_conv = b.conv(sum(_widths(num_channels_up[0], alpha)), num_output_channels, 3 if is3d else 1)
    model.put(out_box)
    if isinstance(last_act_fun, str) and last_act_fun.lower() == "none":
        last_act_fun = None
    if last_act_fun is not None:
        model.put(_act_module(last_act_fun))
    spec = {"is3d": is3d, "act": act_fun, "inputdepth": num_input_channels, "upsample": upsample_mode,
            "first": first_spec, "levels": levels, "out": out_conv, "last_act": last_act_fun,
            "dropout": float(dropout)}
    model._init_runtime(spec, precision)
    return model


def MulResUnet3D(num_input_channels=1, num_output_channels=1, num_channels_down=(16, 32, 64, 128, 256),
                 num_channels_up=(16, 32, 64, 128, 256), num_channels_skip=(16, 32, 64, 128), alpha=1.67,
                 last_act_fun=None, need_bias=True, upsample_mode="nearest", act_fun="LeakyReLU", dropout=0.,
                 precision="fp32"):
    """3-D MultiRes U-Net — same signature as ``architectures/mulresunet.py:188-198`` (+ ``precision``)."""
    return _build(True, num_input_channels, num_output_channels, list(num_channels_down), list(num_channels_up),
                  list(num_channels_skip), alpha, last_act_fun, need_bias, upsample_mode, act_fun, dropout, precision)


def MulResUnet(num_input_channels=1, num_output_channels=1, num_channels_down=(16, 32, 64, 128, 256),
               num_channels_up=(16, 32, 64, 128, 256), num_channels_skip=(16, 32, 64, 128), alpha=1.67,
               last_act_fun=None, need_bias=True, upsample_mode="nearest", act_fun="LeakyReLU", dropout=0.,
               precision="fp32"):
    """2-D MultiRes U-Net — same signature as ``architectures/mulresunet.py:116-126`` (+ ``precision``)."""
    return _build(False, num_input_channels, num_output_channels, list(num_channels_down), list(num_channels_up),
                  list(num_channels_skip), alpha, last_act_fun, need_bias, upsample_mode, act_fun, dropout, precision)


def AttMulResUnet2D(num_input_channels=1, num_output_channels=3, num_channels_down=(16, 32, 64, 128, 256), alpha=1.67,
                    last_act_fun=None, need_bias=True, upsample_mode="nearest", act_fun="LeakyReLU", dropout=0.,
                    precision="fp32"):
    """Attention MultiRes U-Net (2-D) — same signature as ``architectures/attention.py:197-211`` (+ ``precision``).

    Module tree, registration order and names follow ``attention.py:213-247`` (``down_mb<i>``, ``down<i>``,
    ``up_mb<i>``, ``att<i>`` = GridAttentionBlock with ``W_g`` / ``W_x`` / ``psi``, ``up<i>``, ``outconv``), so the
    ``state_dict`` keys, the RNG draws of construction and of ``init_weights`` are the reference's."""
    filters = list(num_channels_down)
    n_scales = len(filters)
    if isinstance(upsample_mode, (list, tuple)):
        if len(set(upsample_mode[1:])) > 1:
            raise NotImplementedError("per-scale upsample modes are not part of the accelerated path")
        upsample_mode = upsample_mode[-1]
    if act_fun not in _ACTS:
        raise NotImplementedError("unknown activation function %r" % (act_fun,))
    b = _Builder(False, act_fun, need_bias, dropout)
    model = DeepPriorNet()
    depths = [num_input_channels]
    down_mb = []
    for i in range(n_scales):
        mrb, mrb_spec = b.block(filters[i], depths[-1], alpha)
        depths.append(mrb.out_dim)
        model.add_module("down_mb%d" % (i + 1), mrb)
        down_mb.append(mrb_spec)
    stages = []
    for i in range(1, n_scales):
        dbox, dconv = b.conv(depths[i], depths[i], 3, stride=2)
        dbn = b.BN(depths[i])
        model.add_module("down%d" % i, nn.Sequential(dbox, dbn, _act_module(act_fun), b.Drop(dropout)))
        up_mb, up_spec = b.block(filters[-(i + 1)], depths[-i] + depths[-(i + 1)], alpha)
        model.add_module("up_mb%d" % i, up_mb)
        # GridAttentionBlock(F_g, F_l, F_int) — attention.py:86-105
        F_g, F_l, F_int = depths[-i], depths[-(i + 1)], filters[-i]
        att = _Holder()
        gbox, gconv = b.conv(F_g, F_int, 1)
        gbn = b.BN(F_int)
        att.add_module("W_g", nn.Sequential(gbox, gbn))
        xbox, xconv = b.conv(F_l, F_int, 3, stride=2)
        xbn = b.BN(F_int)
        att.add_module("W_x", nn.Sequential(xbox, xbn))
        pbox, pconv = b.conv(F_int, 1, 1)
        att.add_module("psi", nn.Sequential(pbox, nn.Sigmoid(), nn.Upsample(scale_factor=2, mode="bilinear")))
        att.add_module("relu", nn.ReLU(inplace=True))
        model.add_module("att%d" % i, att)
        model.add_module("up%d" % i, nn.Upsample(scale_factor=2, mode=upsample_mode))
        stages.append({"down": (dconv, dbn), "dec": up_spec,
                       "att": {"W_g": (gconv, gbn), "W_x": (xconv, xbn), "psi": pconv}})
    if isinstance(last_act_fun, str) and last_act_fun.lower() == "none":
        last_act_fun = None
    out_box, out_conv = b.conv(depths[1], num_output_channels, 1)
    if last_act_fun is not None:
        model.add_module("outconv", nn.Sequential(out_box, _act_module(last_act_fun)))
    else:
        model.add_module("outconv", out_box)
    # stage i of the constructor (down<i>, shallow -> deep) pairs with decoder up_mb<n-i> / att<n-i> in forward()
    # (attention.py:249-262): level j (1 = shallowest) = down<j>, down_mb<j+1>, att<n-j>, up<n-j>, up_mb<n-j>
    levels = []
    for j in range(1, n_scales):
        levels.append({"down": stages[j - 1]["down"], "enc": down_mb[j], "att": stages[n_scales - j - 1]["att"],
                       "dec": stages[n_scales - j - 1]["dec"]})
    spec = {"kind": "attmultiunet", "is3d": False, "act": act_fun, "inputdepth": num_input_channels,
            "upsample": upsample_mode, "first": down_mb[0], "levels": levels, "out": out_conv,
            "last_act": last_act_fun, "dropout": float(dropout)}
    model._init_runtime(spec, precision)
    return model


def get_net(args, outchannel=1):
    """``architectures.get_net`` (architectures/__init__.py:10-86) for the hot-path architecture ``multiunet`` and its
    2-D attention variant ``attmultiunet``.

    ``--net load`` resolves to the multiunet constructor in the reference (its ``else`` branches), so it does here
    too; the other choices are outside the accelerated path.  Like the reference,
    ``attmultiunet`` only exists for ``--datadim 2d / 2.5d``; with ``3d`` the reference falls through to MulResUnet3D."""
    net = getattr(args, "net", "multiunet")
    is2d = args.datadim in ("2d", "2.5d")
    if net == "attmultiunet" and is2d:
        return AttMulResUnet2D(num_input_channels=args.inputdepth, num_output_channels=outchannel,
                               num_channels_down=args.filters, upsample_mode=args.upsample, need_bias=True,
                               act_fun=args.activation, last_act_fun=args.last_activation, dropout=args.dropout,
                               precision=getattr(args, "precision", "fp32"))
    if net not in ("multiunet", "load", "attmultiunet"):
        raise NotImplementedError("--net %s is outside the B200 hot path (only multiunet / attmultiunet / load)" % args.net)
    ctor = MulResUnet if is2d else MulResUnet3D
    return ctor(num_input_channels=args.inputdepth, num_output_channels=outchannel, num_channels_down=args.filters,
                num_channels_up=args.filters, num_channels_skip=args.skip, upsample_mode=args.upsample,
                need_bias=True, act_fun=args.activation, last_act_fun=args.last_activation, dropout=args.dropout,
                precision=getattr(args, "precision", "fp32"))
