"""Plan builder + executor of the deep-prior hot path on one B200.

The reference evaluates ``net(input_)``, the masked loss, ``backward()`` and ``Adam.step()`` as ~1500
separate PyTorch operator calls per iteration (``main.py:141-220``).  Here the static graph
of the MultiRes U-Net is compiled ONCE per (network, patch shape) into a flat list of C-ABI kernel launches
over preallocated channels-last buffers:

    pack weights -> forward ops -> masked loss (+metrics, +dL/dout) -> backward ops (reverse) -> Adam -> tick

All pointers are fixed after the build, so a whole iteration can be captured in a CUDA graph and replayed
with no host work (`Engine.capture`).  There is no autograd here: every op carries its hand-written
backward.  PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import ConvGeom, lib
from .layout import ChannelLayout, pad4

_VP = C.c_void_p


def _vp(x) -> C.c_void_p:
    return _VP(int(x) if x else None)


class _Arena:
    """Bump allocator over a few large zero-filled device chunks.  A compiled plan owns ~600 buffers; one ``torch.zeros``
    each means ~600 fill kernels per engine build (they were half of the launches of a small run).  Buffers of at least
    a quarter chunk get an allocation of their own; every address is 256-byte aligned (TMA needs 16)."""
    CHUNK = 64 << 20

    def __init__(self, device):
        self.device, self.chunks, self.cur, self.off = device, [], None, 0

    def alloc(self, nbytes: int) -> torch.Tensor:
        nbytes = max((int(nbytes) + 255) & ~255, 256)
        if nbytes >= self.CHUNK // 4:
            t = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            self.chunks.append(t)
            return t
        if self.cur is None or self.off + nbytes > self.CHUNK:
            self.cur = torch.zeros(self.CHUNK, dtype=torch.uint8, device=self.device)
            self.chunks.append(self.cur)
            self.off = 0
        v = self.cur[self.off:self.off + nbytes]
        self.off += nbytes
        return v

    def f32(self, n: int) -> torch.Tensor:
        n = int(n)
        return self.alloc(4 * n).view(torch.float32)[:n]

    def u8(self, n: int) -> torch.Tensor:
        return self.alloc(n)[:int(n)]


class Store:
    """One device allocation ``[nvox][ld]`` and (lazily) its gradient twin."""

    def __init__(self, eng: "Engine", dims: Tuple[int, int, int], ld: int):
        self.dims = dims
        self.nvox = dims[0] * dims[1] * dims[2]
        self.ld = ld
        self.data = eng.arena.f32(self.nvox * ld).view(self.nvox, ld)
        self.grad: Optional[torch.Tensor] = None
        self.goff, self.gld = 0, ld      # the gradient may be a channel slice of a wider buffer (fused data gradients)
        self.eng = eng
        self.writers: List[Tuple[int, int, object, str]] = []   # grad writers in forward order

    def need_grad(self):
        if self.grad is None:
            self.grad = self.eng.arena.f32(self.nvox * self.ld).view(self.nvox, self.ld)

    def grad_into(self, buf: torch.Tensor, off: int):
        """the gradient of this tensor lives in channels [off, off + ld) of `buf` ([nvox][wider pitch])"""
        assert buf.shape[0] == self.nvox and off % 4 == 0 and off + self.ld <= buf.shape[1]
        self.grad, self.goff, self.gld = buf, off, int(buf.shape[1])


class Tn:
    """Channels-last tensor view: a channel range of a Store."""

    def __init__(self, store: Store, coff: int, layout: ChannelLayout, needs_grad: bool = True):
        self.store, self.coff, self.layout, self.needs_grad = store, coff, layout, needs_grad
        self.stats_ws: Optional[torch.Tensor] = None    # set when a producer emits (sum, sumsq) for this tensor
        if needs_grad:
            store.need_grad()

    @property
    def C(self) -> int:
        return self.layout.C_p

    @property
    def ld(self) -> int:
        return self.store.ld

    @property
    def dims(self):
        return self.store.dims

    @property
    def nvox(self) -> int:
        return self.store.nvox

    @property
    def ptr(self) -> int:
        return self.store.data.data_ptr() + 4 * self.coff

    @property
    def gptr(self) -> int:
        return self.store.grad.data_ptr() + 4 * (self.store.goff + self.coff)

    @property
    def gld(self) -> int:
        return self.store.gld

    def slice(self, coff: int, layout: ChannelLayout) -> "Tn":
        assert coff % 4 == 0 and coff + layout.C_p <= self.C
        return Tn(self.store, self.coff + coff, layout, self.needs_grad)

    def view(self) -> torch.Tensor:
        return self.store.data[:, self.coff:self.coff + self.C]

    def gview(self) -> torch.Tensor:
        o = self.store.goff + self.coff
        return self.store.grad[:, o:o + self.C]


class FlatParams:
    """All parameters in one flat fp32 buffer (+ one flat gradient buffer); ``p.data`` / ``p.grad`` become views.

    BatchNorm running statistics get the same treatment so that kernels can update them in place."""

    def __init__(self, net: torch.nn.Module, device):
        self.net, self.device = net, device
        self.plist = list(net.parameters())
        self.blist = [b for _, b in net.named_buffers() if b.dtype == torch.float32]
        self.ilist = [b for _, b in net.named_buffers() if b.dtype == torch.int64]
        self.poff, n = [], 0
        for p in self.plist:
            self.poff.append(n)
            n += pad4(p.numel())
        self.n = max(n, 4)
        self.boff, nb = [], 0
        for b in self.blist:
            self.boff.append(nb)
            nb += pad4(b.numel())
        self.P = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.G = torch.zeros(self.n, dtype=torch.float32, device=device)
        self.B = torch.zeros(max(nb, 4), dtype=torch.float32, device=device)
        self.I = torch.zeros(max(len(self.ilist), 1), dtype=torch.int64, device=device)
        self._index = {id(p): i for i, p in enumerate(self.plist)}
        self._bindex = {id(b): i for i, b in enumerate(self.blist)}
        self._iindex = {id(b): i for i, b in enumerate(self.ilist)}
        self.adopt()

    def rebind(self, net: torch.nn.Module) -> bool:
        """Point the flat buffers at another network instance of the same architecture (a fresh network per
        patch, main.py:286-290) so the compiled plan and its CUDA graph are reused.  False if incompatible."""
        plist = list(net.parameters())
        blist = [b for _, b in net.named_buffers() if b.dtype == torch.float32]
        ilist = [b for _, b in net.named_buffers() if b.dtype == torch.int64]
        if [tuple(p.shape) for p in plist] != [tuple(p.shape) for p in self.plist] or \
                [tuple(b.shape) for b in blist] != [tuple(b.shape) for b in self.blist] or len(ilist) != len(self.ilist):
            return False
        for p in self.plist:       # detach the old network from the flat storage
            p.data = p.data.clone()
            p.grad = None
        for b in self.blist + self.ilist:
            b.data = b.data.clone()
        self.net, self.plist, self.blist, self.ilist = net, plist, blist, ilist
        self._index = {id(p): i for i, p in enumerate(self.plist)}
        self._bindex = {id(b): i for i, b in enumerate(self.blist)}
        self._iindex = {id(b): i for i, b in enumerate(self.ilist)}
        self.adopt()
        return True

    def adopt(self):
        """Copy the current parameter values in and re-point every tensor at the flat storage."""
        with torch.no_grad():
            for p, off in zip(self.plist, self.poff):
                v = self.P[off:off + p.numel()].view(p.shape)
                if p.data.data_ptr() != v.data_ptr():
                    v.copy_(p.data.to(device=self.device, dtype=torch.float32))
                    p.data = v
            for b, off in zip(self.blist, self.boff):
                v = self.B[off:off + b.numel()].view(b.shape)
                if b.data.data_ptr() != v.data_ptr():
                    v.copy_(b.data.to(device=self.device, dtype=torch.float32))
                    b.data = v
            for i, b in enumerate(self.ilist):
                v = self.I[i:i + 1].view(b.shape)
                if b.data.data_ptr() != v.data_ptr():
                    v.copy_(b.data.to(device=self.device))
                    b.data = v

    def stale(self) -> bool:
        if not self.plist:
            return False
        p0, pl = self.plist[0], self.plist[-1]
        return (p0.data.data_ptr() != self.P.data_ptr() + 4 * self.poff[0]
                or pl.data.data_ptr() != self.P.data_ptr() + 4 * self.poff[-1])

    def bind_grads(self):
        for p, off in zip(self.plist, self.poff):
            p.grad = self.G[off:off + p.numel()].view(p.shape)

    def ptr(self, p) -> int:
        return self.P.data_ptr() + 4 * self.poff[self._index[id(p)]]

    def gptr(self, p) -> int:
        return self.G.data_ptr() + 4 * self.poff[self._index[id(p)]]

    def bptr(self, b) -> int:
        return self.B.data_ptr() + 4 * self.boff[self._bindex[id(b)]]

    def iptr(self, b) -> int:
        return self.I.data_ptr() + 8 * self._iindex[id(b)]


_SKIP_CALLS = frozenset(n for n in os.environ.get("DPI_TIMING_SKIP_CALLS", "").split(",") if n)


# DPI_FUSE_NEXT_REDUCE=0: every BatchNorm-backward reduce is its own launch (A/B switch of dpi_bn_next_reduce)
_FUSE_NEXT_REDUCE = os.environ.get("DPI_FUSE_NEXT_REDUCE", "1") != "0"
# DPI_FUSE_PART_STATS=1: the BatchNorm over a block's concatenated branches reads the statistics its three producers
# leave (dpi_bn_finalize_parts) instead of running a statistics pass of its own over the concatenation.  Measured
# (256,128,128): 25.84 -> 25.78 ms per iteration, 64^3: 4.37 -> 4.44 ms - the three narrow BatchNorm + activation passes
# (4 / 8 / 16 channels) lose in statistics mode (256-thread CTAs, fp64 sums, a row flush each) what the removed pass
# saves, so it is OFF by default.
_FUSE_PART_STATS = os.environ.get("DPI_FUSE_PART_STATS", "0") == "1"
# small tensors are launch-bound and the fused kernels (256-thread CTAs that also write statistics rows) cost more than the
# launch they save: measured on a 64^3 patch, fusing at every level +1.3 % per iteration
_FUSE_NEXT_MIN_VOX = int(os.environ.get("DPI_FUSE_NEXT_MIN_VOX", "131072"))


class _Call:
    """A pre-marshalled C-ABI call; the stream is appended at run time."""
    __slots__ = ("fn", "args", "name", "lane")

    def __init__(self, name: str, *args):
        self.lane = 0
        fn = getattr(lib, name)
        ats = fn.argtypes[:-1]
        assert len(ats) == len(args), (name, len(ats), len(args))
        conv = []
        for at, a in zip(ats, args):
            if at is _VP:
                conv.append(_vp(a))
            elif isinstance(a, (C.Structure, C._Pointer)) or hasattr(a, "_obj"):
                conv.append(a)
            else:
                conv.append(at(a))
        self.fn, self.args, self.name = fn, tuple(conv), name

    def __call__(self, st):
        if self.name in _SKIP_CALLS:       # DPI_TIMING_SKIP_CALLS: what-if timing experiments only, results are garbage
            return
        rc = self.fn(*self.args, st)
        if rc != 0:
            raise _lib.DpiError("%s failed (rc=%d): %s" % (self.name, rc, _lib.last_error()))


class _SideCall:
    """A call that may run on the engine's side stream, concurrently with what follows it on the main stream.

    Used for the weight gradients: dw is only needed by the un-pack at the end of the backward pass, while the tensor-
    bound, persistent wgrad kernels (one 200 KB-SMEM CTA per SM) leave room on every SM for the HBM-bound BatchNorm /
    activation streams of the main chain."""
    __slots__ = ("calls", "lane")

    def __init__(self, calls):
        self.calls = list(calls)
        self.lane = 0


class _Wait:
    """Stream dependency: everything issued so far on lane `signaler` happens before what follows on lane `waiter`."""
    __slots__ = ("waiter", "signaler")

    def __init__(self, waiter: int, signaler: int):
        self.waiter, self.signaler = waiter, signaler


class Op:
    lane = 0                    # 0 = main chain; 1 = the shortcut branch of a MultiRes block (runs concurrently)

    def emit_pack(self) -> List[_Call]:
        return []

    def emit_fwd(self) -> List[_Call]:
        return []

    def emit_bwd(self) -> List[_Call]:
        return []


class MarkerOp(Op):
    """Fork / join point between lanes (no kernel): a _Wait in the forward and / or the backward launch list."""

    def __init__(self, fwd=None, bwd=None):
        self.fwd, self.bwd = fwd, bwd

    def emit_fwd(self):
        return [_Wait(*self.fwd)] if self.fwd else []

    def emit_bwd(self):
        return [_Wait(*self.bwd)] if self.bwd else []


class ConvOp(Op):
    """nn.Conv3d / nn.Conv2d forward + dgrad + wgrad (base.py:117-126,169-180)."""

    def __init__(self, eng: "Engine", x: Tn, conv: torch.nn.Module, out_layout: ChannelLayout, bn_follows: bool):
        w = conv.weight
        k = tuple(w.shape[2:])
        if len(k) == 2:
            k = (1,) + k
        self.eng, self.x, self.conv, self.bn_follows = eng, x, conv, bn_follows
        self.Cout_l, self.Cin_l = int(w.shape[0]), int(w.shape[1])
        assert x.layout.C_l == self.Cin_l and out_layout.C_l == self.Cout_l, "channel mismatch"
        self.taps = k[0] * k[1] * k[2]
        stride = int(conv.stride[0])
        D, H, W = x.dims
        self.geom = ConvGeom(D, H, W, x.C, out_layout.C_p, k[0], k[1], k[2], stride)

        def osz(n, kk):
            s = stride if kk > 1 else 1
            p = (kk - 1) // 2
            return (n + 2 * p - kk) // s + 1

        odims = (osz(D, k[0]), osz(H, k[1]), osz(W, k[2]))
        self.y = eng.new_tensor(odims, out_layout)
        if bn_follows and eng.prec == _lib.PREC_TF32 and os.environ.get("DPI_CONV_STATS", "1") != "0":
            # the conv leaves the BatchNorm partial sums of y in this workspace (fused into the tcgen05 epilogue where
            # the kernel supports it), so the BnActOp that follows skips its own statistics pass over y
            self.y.stats_ws = eng.stats_ws(out_layout.C_p)
        nw = out_layout.C_p * self.taps * x.C
        self.wf = eng.zeros(nw)
        self.wd = eng.zeros(nw) if x.needs_grad else None
        self.dwp = eng.zeros(nw)
        self.bp = eng.zeros(out_layout.C_p)
        self.cout_map, self.cin_map = eng.map_tensor(out_layout), eng.map_tensor(x.layout)
        eng.wgrad_ws_bytes = max(eng.wgrad_ws_bytes, int(lib.dpi_conv_wgrad_workspace_bytes(C.byref(self.geom))))
        eng.max_C = max(eng.max_C, out_layout.C_p)
        self.acc = {"dx": False}
        self.bwd_pre_wait = None        # (waiter, signaler): x.grad was first written on another lane
        self.fused: Optional["ConvOp"] = None      # the 1x1 conv whose data gradient this op computes along with its own
        self.dgrad_delegated = False               # ... and, on that conv: the partner writes x.grad for both
        self.cat = None                            # (buffer, Cc, channel offset, tap0, taps) of the fused weight pack
        if x.needs_grad:
            eng.register_grad_write(x, self, "dx")

    def fuse_dgrad_with(self, other: "ConvOp") -> bool:
        """Fused data gradient with the 1x1 conv `other` that reads the same x (Block.shortcut / ResPath.conv1x1): both
        output gradients become channel slices of ONE buffer, the transposed weights one [Cin][taps][Ca + Cb] pack (1x1
        at the centre tap), and one march dgrad launch writes x.grad for both - x.grad is written once instead of
        written and then read-modify-written.  False (nothing changed) where that launch would not run on the kernels
        that skip the zero blocks."""
        eng, x = self.eng, self.x
        if not x.needs_grad or other.x is not x or other.taps != 1 or self.taps not in (9, 27) or eng.prec != _lib.PREC_TF32:
            return False
        if int(self.conv.stride[0]) != 1 or os.environ.get("DPI_FUSED_DGRAD", "1") == "0":
            return False
        Cc = self.y.C + other.y.C
        g = self.geom
        geom_cat = ConvGeom(g.D, g.H, g.W, x.C, Cc, g.kd, g.kh, g.kw, 1)
        if not lib.dpi_conv_dgrad_fused_supported(C.byref(geom_cat), self.y.C):
            return False
        buf = eng.arena.f32(self.y.nvox * Cc).view(self.y.nvox, Cc)
        self.y.store.grad_into(buf, 0)
        other.y.store.grad_into(buf, self.y.C)
        self.geom_cat = geom_cat
        self.wd_cat = eng.zeros(x.C * self.taps * Cc)
        self.cat = (self.wd_cat, Cc, 0, 0, self.taps)
        other.cat = (self.wd_cat, Cc, self.y.C, self.taps // 2, self.taps)
        self.fused, other.dgrad_delegated = other, True
        x.store.writers = [w for w in x.store.writers if not (w[2] is other and w[3] == "dx")]
        return True

    def _pk(self):
        return (self.cout_map.data_ptr(), self.cin_map.data_ptr(), self.Cout_l, self.Cin_l, self.y.C, self.x.C,
                self.taps)

    def pack_job(self) -> "_lib.PackJob":
        """this layer's entry of the batched pack / unpack job table (dpi_pack_job)"""
        P = self.eng.params
        bias = self.conv.bias
        cat = self.cat if self.cat is not None else (None, 0, 0, 0, 0)
        return _lib.PackJob(P.ptr(self.conv.weight), P.ptr(bias) if bias is not None else None, self.wf.data_ptr(),
                            self.wd.data_ptr() if self.wd is not None else None, self.bp.data_ptr(),
                            self.dwp.data_ptr(), P.gptr(self.conv.weight), self.cout_map.data_ptr(),
                            self.cin_map.data_ptr(), self.Cout_l, self.Cin_l, self.y.C, self.x.C, self.taps, 0,
                            cat[0].data_ptr() if cat[0] is not None else None, cat[1], cat[2], cat[4], cat[3])

    def emit_fwd(self):
        x, y = self.x, self.y
        if y.stats_ws is not None:
            return [_Call("dpi_conv_fwd_stats", x.ptr, x.ld, self.wf.data_ptr(), self.bp.data_ptr(), y.ptr, y.ld,
                          C.byref(self.geom), self.eng.prec, y.stats_ws.data_ptr())]
        return [_Call("dpi_conv_fwd", x.ptr, x.ld, self.wf.data_ptr(), self.bp.data_ptr(), y.ptr, y.ld,
                      C.byref(self.geom), self.eng.prec)]

    def emit_bwd(self):
        eng, x, y, P = self.eng, self.x, self.y, self.eng.params
        calls = []
        if self.conv.bias is not None and not self.bn_follows:
            # a bias that feeds a BatchNorm has an exactly-zero gradient (the batch mean absorbs it);
            # it is left at 0 instead of reproducing the reference's rounding noise
            ws = eng.bwd_ws_for(self.lane)
            calls.append(_Call("dpi_bias_grad", y.gptr, y.gld, y.nvox, y.C, self.cout_map.data_ptr(),
                               P.gptr(self.conv.bias), ws.data_ptr(), ws.numel()))
        if x.needs_grad and not self.dgrad_delegated:
            if self.bwd_pre_wait is not None:
                calls.append(_Wait(*self.bwd_pre_wait))
            if self.fused is not None:
                # one pass for this conv AND the 1x1 conv that shares its input (their dy sit side by side in one buffer)
                calls.append(_Call("dpi_conv_dgrad_fused", y.gptr, y.gld, self.wd_cat.data_ptr(), x.gptr, x.gld,
                                   C.byref(self.geom_cat), y.C, 1 if self.acc["dx"] else 0, eng.prec))
            else:
                calls.append(_Call("dpi_conv_dgrad", y.gptr, y.gld, self.wd.data_ptr(), x.gptr, x.gld, C.byref(self.geom),
                                   1 if self.acc["dx"] else 0, eng.prec))
        # the weight gradient goes LAST and to the side stream: the data gradient (also a persistent, SMEM-filling
        # kernel) has run by then, so what the wgrad overlaps with on the main stream is the HBM-bound BatchNorm
        # backward of the next unit
        calls.append(_SideCall([_Call("dpi_conv_wgrad", x.ptr, x.ld, y.gptr, y.gld, self.dwp.data_ptr(),
                                      C.byref(self.geom), eng.wgrad_ws.data_ptr(), eng.wgrad_ws.numel() * 4, eng.prec)]))
        return calls


class BnActOp(Op):
    """[training-mode BatchNorm] + [activation] as one streaming pass (base.py:162-166,211-216)."""

    def __init__(self, eng: "Engine", x: Tn, bn: Optional[torch.nn.Module], act: Optional[str],
                 out: Optional[Tn] = None, emit_stats: bool = False, round_out: bool = False):
        self.eng, self.x, self.bn, self.act = eng, x, bn, _lib.ACT_CODES[act]
        tf32 = eng.prec == _lib.PREC_TF32
        self.rf = _lib.ROUND_TF32 if (tf32 and round_out) else 0     # output feeds a tcgen05 conv
        self.rb = _lib.ROUND_TF32 if tf32 else 0                     # x.grad is the dy operand of a conv
        self.out = out if out is not None else eng.new_tensor(x.dims, x.layout)
        assert self.out.C == x.C and self.out.nvox == x.nvox
        self.map = eng.map_tensor(x.layout)
        self.aux = eng.zeros(6 * x.C)
        if bn is not None:
            assert bn.num_features == x.layout.C_l
            self.own_stats = x.stats_ws is None
            self.ws = x.stats_ws if x.stats_ws is not None else eng.stats_ws(x.C)
        if emit_stats:
            self.out.stats_ws = eng.stats_ws(x.C)
        self.acc = {"dx": False}
        eng.register_grad_write(x, self, "dx")
        eng.max_C = max(eng.max_C, x.C)
        # fused BatchNorm-backward reduces (dpi_bn_next_reduce): `next_add` = the residual add whose output this unit
        # normalises (norm2 of a MultiRes block): this unit's apply pass also takes the add's sums; `reduce_fused` = set by
        # the producer of this unit's incoming gradient when it has taken THIS unit's sums
        self.next_add: Optional["AddActOp"] = None
        self.reduce_fused = False

    def _aux(self, i):
        return self.aux.data_ptr() + 4 * i * self.x.C

    def emit_fwd(self):
        x, o, P, bn = self.x, self.out, self.eng.params, self.bn
        ows = o.stats_ws.data_ptr() if o.stats_ws is not None else 0
        if bn is None:
            return [_Call("dpi_affine_act", x.ptr, x.ld, 0, 0, 0, self.act | self.rf, o.ptr, o.ld, x.nvox, x.C, ows)]
        calls = []
        if self.own_stats:
            calls.append(_Call("dpi_channel_stats", x.ptr, x.ld, x.nvox, x.C, self.ws.data_ptr()))
        calls.append(_Call("dpi_bn_finalize", self.ws.data_ptr(), x.nvox, x.C, self.map.data_ptr(), P.ptr(bn.weight),
                           P.ptr(bn.bias), P.bptr(bn.running_mean), P.bptr(bn.running_var),
                           P.iptr(bn.num_batches_tracked), float(bn.momentum), float(bn.eps), self._aux(0),
                           self._aux(1), self._aux(2), self._aux(3)))
        calls.append(_Call("dpi_affine_act", x.ptr, x.ld, self._aux(0), self._aux(2), self._aux(3), self.act | self.rf,
                           o.ptr, o.ld, x.nvox, x.C, ows))
        return calls

    def emit_bwd(self):
        x, o, P, bn, eng = self.x, self.out, self.eng.params, self.bn, self.eng
        acc = 1 if self.acc["dx"] else 0
        optr = o.ptr if self.act else 0
        if bn is None:
            return [_Call("dpi_act_bwd", o.gptr, o.gld, optr, o.ld, self.act | self.rb, x.gptr, x.gld, x.nvox, x.C, acc)]
        # out = act((x - mean) * scale + shift) is re-derived from x inside the kernels (out pointer NULL, scale and
        # shift given): one tensor read less in each of the two passes
        sc, sh = (self._aux(2), self._aux(3)) if self.act else (0, 0)
        ws = eng.bwd_ws_for(self.lane)
        calls = []
        if not self.reduce_fused:
            calls.append(_Call("dpi_bn_bwd_reduce", o.gptr, o.gld, 0, o.ld, self.act, x.ptr, x.ld, self._aux(0),
                               self._aux(1), sc, sh, x.nvox, x.C, ws.data_ptr()))
        calls.append(_Call("dpi_bn_bwd_finalize", ws.data_ptr(), x.nvox, x.C, self.map.data_ptr(), P.gptr(bn.weight),
                           P.gptr(bn.bias), self._aux(4), self._aux(5)))
        add = self.next_add
        if (add is not None and _FUSE_NEXT_REDUCE and x.nvox >= _FUSE_NEXT_MIN_VOX and not acc and add.bn is not None
                and add.multi and add.out is x and add.lane == self.lane):
            # this unit's x is the add's activation output and x.grad, written here only, is the add's incoming gradient:
            # the apply pass also accumulates the add's BatchNorm-backward sums (kind 1)
            nx = _lib.NextReduce.make(1, add.act, add._parts(), add._aux(0), add._aux(1), 0, 0,
                                      eng.bwd_ws_for(add.lane).data_ptr())
            calls.append(_Call("dpi_bn_bwd_apply_next", o.gptr, o.gld, 0, o.ld, self.act | self.rb, x.ptr, x.ld,
                               self._aux(0), self._aux(1), self._aux(2), sh, self._aux(4), self._aux(5), x.gptr, x.gld,
                               x.nvox, x.C, acc, nx))
            add.reduce_fused = True
        elif (add is not None and _FUSE_NEXT_REDUCE and not acc and add.bn is None and not add.multi and add.out is x
              and add.lane == self.lane and not add.acc["dp"] and not add.acc["dq0"]):
            # the add has no BatchNorm (ResPath): g = x.grad * act'(x) is the gradient of both its addends - written
            # here, straight into q.grad and p.grad (kind 3); the add emits no backward launches of its own
            p_, q_ = add.p, add.qs[0]
            nx = _lib.NextReduce.make(3, add.act, _lib.Parts.make([p_.gptr], [p_.gld], [p_.C]), 0, 0, 0, 0, 0)
            calls.append(_Call("dpi_bn_bwd_apply_next", o.gptr, o.gld, 0, o.ld, self.act | self.rb, x.ptr, x.ld,
                               self._aux(0), self._aux(1), self._aux(2), sh, self._aux(4), self._aux(5), q_.gptr, q_.gld,
                               x.nvox, x.C, 0, nx))
            add.bwd_fused = True
        else:
            calls.append(_Call("dpi_bn_bwd_apply", o.gptr, o.gld, 0, o.ld, self.act | self.rb, x.ptr, x.ld, self._aux(0),
                               self._aux(1), self._aux(2), sh, self._aux(4), self._aux(5), x.gptr, x.gld, x.nvox, x.C,
                               acc))
        return calls


class AddActOp(Op):
    """out = act(p + [BN](q))  — the residual adds of Block*/ResPath* (mulresunet.py:33-34,60,90-93,109-110).

    ``q`` may be a LIST of tensors: the branch outputs of a MultiRes block, which the reference concatenates
    (``torch.cat``, mulresunet.py:31,89).  They stay in separate dense buffers and the kernels address them as the
    channel parts of one tensor (``dpi_parts``); no concat buffer exists."""

    def __init__(self, eng: "Engine", p: Tn, q, bn_q: Optional[torch.nn.Module], act: Optional[str],
                 emit_stats: bool = False, round_out: bool = False, q_layout: Optional[ChannelLayout] = None):
        self.eng, self.p, self.bn, self.act = eng, p, bn_q, _lib.ACT_CODES[act]
        self.qs: List[Tn] = list(q) if isinstance(q, (list, tuple)) else [q]
        self.multi = len(self.qs) > 1
        self.qlay = q_layout if q_layout is not None else self.qs[0].layout
        self.C = self.qlay.C_p
        self.nvox = self.qs[0].nvox
        self.dims = self.qs[0].dims
        self.rf = _lib.ROUND_TF32 if (eng.prec == _lib.PREC_TF32 and round_out) else 0
        assert p.C == self.C and p.nvox == self.nvox and sum(t.C for t in self.qs) == self.C
        assert all(t.nvox == self.nvox for t in self.qs) and len(self.qs) <= 4
        self.out = eng.new_tensor(self.dims, self.qlay)
        self.map = eng.map_tensor(self.qlay)
        self.aux = eng.zeros(6 * self.C)
        if bn_q is not None:
            assert bn_q.num_features == self.qlay.C_l
            # every branch output already carries the statistics its producer left (dpi_bn_finalize_parts): no pass of
            # its own over the concatenation
            self.part_stats = self.multi and _FUSE_PART_STATS and all(t.stats_ws is not None for t in self.qs)
            self.own_stats = (self.multi and not self.part_stats) or (not self.multi and self.qs[0].stats_ws is None)
            self.ws = eng.stats_ws(self.C) if self.own_stats else self.qs[0].stats_ws
        if emit_stats:
            self.out.stats_ws = eng.stats_ws(self.C)
        self.acc = {"dp": False}
        self.reduce_fused = False                     # set by the BnActOp that follows (its apply took this op's sums)
        self.bwd_fused = False                        # ... or this op's whole backward (no BatchNorm here: kind 3)
        self.next_bn: Optional[BnActOp] = None        # the unit that produced p (the block's shortcut conv + BN)
        eng.register_grad_write(p, self, "dp")
        for i, t in enumerate(self.qs):
            self.acc["dq%d" % i] = False
            eng.register_grad_write(t, self, "dq%d" % i)
        eng.max_C = max(eng.max_C, self.C)

    def _aux(self, i):
        return self.aux.data_ptr() + 4 * i * self.C

    def _parts(self, grad: bool = False) -> "_lib.Parts":
        return _lib.Parts.make([t.gptr if grad else t.ptr for t in self.qs], [t.gld if grad else t.ld for t in self.qs],
                               [t.C for t in self.qs])

    def emit_fwd(self):
        p, o, P, bn = self.p, self.out, self.eng.params, self.bn
        q = self.qs[0]
        ows = o.stats_ws.data_ptr() if o.stats_ws is not None else 0
        if bn is None:
            if self.multi:
                return [_Call("dpi_add_affine_act_parts", p.ptr, p.ld, self._parts(), 0, 0, 0, self.act | self.rf, o.ptr,
                              o.ld, self.nvox, self.C, ows)]
            return [_Call("dpi_add_affine_act", p.ptr, p.ld, q.ptr, q.ld, 0, 0, 0, self.act | self.rf, o.ptr, o.ld,
                          self.nvox, self.C, ows)]
        calls = []
        if self.own_stats:
            if self.multi:
                calls.append(_Call("dpi_channel_stats_parts", self._parts(), self.nvox, self.C, self.ws.data_ptr()))
            else:
                calls.append(_Call("dpi_channel_stats", q.ptr, q.ld, self.nvox, self.C, self.ws.data_ptr()))
        if self.multi and self.part_stats:
            sp = _lib.StatsParts.make([t.stats_ws.data_ptr() for t in self.qs], [t.C for t in self.qs])
            calls.append(_Call("dpi_bn_finalize_parts", sp, self.nvox, self.C, self.map.data_ptr(), P.ptr(bn.weight),
                               P.ptr(bn.bias), P.bptr(bn.running_mean), P.bptr(bn.running_var),
                               P.iptr(bn.num_batches_tracked), float(bn.momentum), float(bn.eps), self._aux(0),
                               self._aux(1), self._aux(2), self._aux(3)))
        else:
            calls.append(_Call("dpi_bn_finalize", self.ws.data_ptr(), self.nvox, self.C, self.map.data_ptr(),
                               P.ptr(bn.weight), P.ptr(bn.bias), P.bptr(bn.running_mean), P.bptr(bn.running_var),
                               P.iptr(bn.num_batches_tracked), float(bn.momentum), float(bn.eps), self._aux(0),
                               self._aux(1), self._aux(2), self._aux(3)))
        if self.multi:
            calls.append(_Call("dpi_add_affine_act_parts", p.ptr, p.ld, self._parts(), self._aux(0), self._aux(2),
                               self._aux(3), self.act | self.rf, o.ptr, o.ld, self.nvox, self.C, ows))
        else:
            calls.append(_Call("dpi_add_affine_act", p.ptr, p.ld, q.ptr, q.ld, self._aux(0), self._aux(2), self._aux(3),
                               self.act | self.rf, o.ptr, o.ld, self.nvox, self.C, ows))
        return calls

    def emit_bwd(self):
        p, o, P, bn, eng = self.p, self.out, self.eng.params, self.bn, self.eng
        q = self.qs[0]
        optr = o.ptr if self.act else 0
        # p.grad = dy * act'(out): a pass of its own, unless the multi-part BN-backward apply below can emit it as a
        # second output (p.grad not accumulated into: the shortcut branch has no other consumer)
        if self.bwd_fused:
            return []
        fuse_dp = bn is not None and self.multi and not self.acc["dp"] and os.environ.get("DPI_FUSE_DP", "1") != "0"
        calls = [] if fuse_dp else [_Call("dpi_act_bwd", o.gptr, o.gld, optr, o.ld, self.act, p.gptr, p.gld, self.nvox,
                                          self.C, 1 if self.acc["dp"] else 0)]
        if bn is None:
            off = 0
            for i, t in enumerate(self.qs):       # one channel slice of dy / out per part
                calls.append(_Call("dpi_act_bwd", o.gptr + 4 * off, o.gld, (optr + 4 * off) if optr else 0, o.ld, self.act,
                                   t.gptr, t.gld, self.nvox, t.C, 1 if self.acc["dq%d" % i] else 0))
                off += t.C
            return calls
        if self.multi:
            mask = sum((1 << i) for i in range(len(self.qs)) if self.acc["dq%d" % i])
            if not self.reduce_fused:
                calls.append(_Call("dpi_bn_bwd_reduce_parts", o.gptr, o.gld, optr, o.ld, self.act, self._parts(),
                                   self._aux(0), self._aux(1), self.nvox, self.C, eng.bwd_ws_for(self.lane).data_ptr()))
            calls.append(_Call("dpi_bn_bwd_finalize", eng.bwd_ws_for(self.lane).data_ptr(), self.nvox, self.C,
                               self.map.data_ptr(), P.gptr(bn.weight), P.gptr(bn.bias), self._aux(4), self._aux(5)))
            sb = self.next_bn
            if (sb is not None and _FUSE_NEXT_REDUCE and self.nvox >= _FUSE_NEXT_MIN_VOX and fuse_dp and optr
                    and sb.bn is not None and sb.out is p and sb.x.C == self.C):
                # p.grad, the second output of this pass, is the incoming gradient of the unit that produced p (the
                # block's shortcut conv + BN): its BatchNorm-backward sums are taken here too (kind 2)
                sbsc, sbsh = (sb._aux(2), sb._aux(3)) if sb.act else (0, 0)
                nx = _lib.NextReduce.make(2, sb.act, _lib.Parts.make([sb.x.ptr], [sb.x.ld], [sb.x.C]), sb._aux(0),
                                          sb._aux(1), sbsc, sbsh, eng.bwd_ws_for(sb.lane).data_ptr())
                calls.append(_Call("dpi_bn_bwd_apply_parts_next", o.gptr, o.gld, optr, o.ld, self.act, self._parts(),
                                   self._aux(0), self._aux(1), self._aux(2), self._aux(4), self._aux(5),
                                   self._parts(grad=True), mask, p.gptr, p.gld, self.nvox, self.C, nx))
                sb.reduce_fused = True
            else:
                calls.append(_Call("dpi_bn_bwd_apply_parts", o.gptr, o.gld, optr, o.ld, self.act, self._parts(),
                                   self._aux(0), self._aux(1), self._aux(2), self._aux(4), self._aux(5),
                                   self._parts(grad=True), mask, p.gptr if fuse_dp else 0, p.gld, self.nvox, self.C))
            return calls
        accq = 1 if self.acc["dq0"] else 0
        calls += [
            _Call("dpi_bn_bwd_reduce", o.gptr, o.gld, optr, o.ld, self.act, q.ptr, q.ld, self._aux(0), self._aux(1),
                  0, 0, self.nvox, self.C, eng.bwd_ws_for(self.lane).data_ptr()),
            _Call("dpi_bn_bwd_finalize", eng.bwd_ws_for(self.lane).data_ptr(), self.nvox, self.C, self.map.data_ptr(), P.gptr(bn.weight),
                  P.gptr(bn.bias), self._aux(4), self._aux(5)),
            _Call("dpi_bn_bwd_apply", o.gptr, o.gld, optr, o.ld, self.act, q.ptr, q.ld, self._aux(0), self._aux(1),
                  self._aux(2), 0, self._aux(4), self._aux(5), q.gptr, q.gld, self.nvox, self.C, accq),
        ]
        return calls


class UpsampleOp(Op):
    """nn.Upsample(scale_factor=2) written straight into the skip-concat slice (mulresunet.py:168,242;
    crop of base.py:302-319,342-357)."""

    def __init__(self, eng: "Engine", x: Tn, out: Tn, mode: str, up_d: bool, feeds_conv: bool = True):
        self.eng, self.x, self.out = eng, x, out
        assert out.C == x.C
        self.mode = _lib.UP_NEAREST if mode == "nearest" else _lib.UP_LINEAR
        # feeds the decoder convs: stored values rounded to TF32
        self.mode_fwd = self.mode | (_lib.ROUND_TF32 if (eng.prec == _lib.PREC_TF32 and feeds_conv) else 0)
        self.up_d = 1 if up_d else 0
        self.acc = {"dx": False}
        eng.register_grad_write(x, self, "dx")

    def emit_fwd(self):
        x, o = self.x, self.out
        return [_Call("dpi_upsample2x_fwd", x.ptr, x.ld, *x.dims, o.ptr, o.ld, *o.dims, x.C, self.mode_fwd, self.up_d)]

    def emit_bwd(self):
        x, o = self.x, self.out
        return [_Call("dpi_upsample2x_bwd", o.gptr, o.gld, *o.dims, x.gptr, x.gld, *x.dims, x.C, self.mode, self.up_d,
                      1 if self.acc["dx"] else 0)]


class GateMulOp(Op):
    """``x * psi`` of GridAttentionBlock.forward (attention.py:107-113): the one-channel attention map (up-sampled to
    the resolution of the skip tensor) scales every channel of x; written straight into the decoder's concat slice."""

    def __init__(self, eng: "Engine", x: Tn, psi: Tn, out: Tn):
        self.eng, self.x, self.psi, self.out = eng, x, psi, out
        assert out.C == x.C and out.nvox == x.nvox == psi.nvox and psi.C == 4
        self.rf = _lib.ROUND_TF32 if eng.prec == _lib.PREC_TF32 else 0       # feeds the decoder convs
        self.acc = {"dx": False, "dpsi": False}
        eng.register_grad_write(x, self, "dx")
        eng.register_grad_write(psi, self, "dpsi")

    def emit_fwd(self):
        x, s, o = self.x, self.psi, self.out
        return [_Call("dpi_gate_mul_fwd", x.ptr, x.ld, s.ptr, s.ld, o.ptr, o.ld, x.nvox, x.C, self.rf)]

    def emit_bwd(self):
        x, s, o = self.x, self.psi, self.out
        assert not self.acc["dpsi"], "the up-sampled attention map has a single consumer"
        return [_Call("dpi_gate_mul_bwd", o.gptr, o.gld, x.ptr, x.ld, s.ptr, s.ld, x.gptr, x.gld, s.gptr, s.gld, x.nvox,
                      x.C, 1 if self.acc["dx"] else 0)]


class Row:
    """One batch row of the shared-network mode (BASELINE config 5): a patch with its own fixed noise
    ``z``, target, mask, best output, loss history and noise counter, evaluated through the ONE compiled plan (and the
    one set of weights) of its engine.  Row 0 shares the engine's Adam step / learning-rate cell, so the optimiser
    state advances once per iteration however many rows a rank holds."""

    def __init__(self, eng: "Engine", primary: bool, seed: int):
        n = eng.out.nvox * eng.out.ld
        dev = eng.device
        self.z = eng.arena.f32(eng.z.nvox * eng.z.ld)
        self.img, self.mask, self.best = eng.zeros(n), eng.zeros(n), eng.zeros(n)
        self.scalars = torch.zeros(8, dtype=torch.float64, device=dev)
        self.hyper = eng.hyper if primary else torch.tensor([1e-3, 1.0], dtype=torch.float64, device=dev)
        self.counter = eng.counter if primary else torch.zeros(2, dtype=torch.int64, device=dev)
        self.best_state = torch.zeros(2, dtype=torch.float64, device=dev)
        self.history = torch.zeros((eng.max_iters, 4), dtype=torch.float64, device=dev)
        self.seed = int(seed)
        self.loss_call = eng._make_loss_call(self.img, self.mask, self.scalars)

    def reset(self):
        self.counter.copy_(torch.tensor([0, self.seed], dtype=torch.int64))
        self.best_state.zero_()
        self.history.zero_()


class Engine:
    """Compiled hot path for one network instance and one patch shape."""

    def __init__(self, net, in_dims: Sequence[int], device, precision: str = "fp32", max_iters: int = 4096):
        if not torch.cuda.is_available():
            raise RuntimeError("deep_prior_interpolation_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.net, self.device = net, torch.device(device)
        self.prec = _lib.PREC_TF32 if precision == "tf32" else _lib.PREC_FP32
        self.loss_kind = _lib.LOSS_CODES["mae"]
        dims = tuple(int(d) for d in in_dims)
        self.dims = dims if len(dims) == 3 else (1,) + dims
        self.max_iters = int(max_iters)
        self._maps: Dict[ChannelLayout, torch.Tensor] = {}
        self._keep: List[torch.Tensor] = []
        self.ops: List[Op] = []
        self.stores: List[Store] = []
        self.wgrad_ws_bytes, self.max_C = 16, 4
        self.arena = _Arena(self.device)
        with torch.cuda.device(self.device):
            self.params = FlatParams(net, self.device)
            self._build()
        self.graph = None
        self._graph_sigma = None
        # DPI_SIDE_STREAM=0: strictly linear launch order (A/B switch)
        n_lanes = 1 + max(op.lane for op in self.ops)
        self.side_streams = None if os.environ.get("DPI_SIDE_STREAM", "1") == "0" else \
            [torch.cuda.Stream(self.device) for _ in range(n_lanes)]     # lanes 1.. + the weight-gradient lane

    def rebind(self, net) -> bool:
        """Reuse this compiled plan (buffers, launch lists, CUDA graph) for another instance of the same
        architecture; the ops only hold flat-buffer addresses, which do not change."""
        if net.spec["is3d"] != self.net.spec["is3d"] or net.precision != self.net.precision:
            return False
        for k in ("kind", "act", "upsample", "last_act", "inputdepth"):
            if net.spec.get(k) != self.net.spec.get(k):
                return False
        if not self.params.rebind(net):
            return False
        self.net = net
        return True

    # ---- allocation helpers -------------------------------------------------------------------
    def zeros(self, n: int) -> torch.Tensor:
        return self.arena.f32(max(int(n), 4))

    def stats_ws(self, Cp: int) -> torch.Tensor:
        return self.arena.u8(int(lib.dpi_stats_workspace_bytes(Cp)))

    def bwd_ws_for(self, lane: int) -> torch.Tensor:
        """BatchNorm-backward partial-sum workspace of a lane (lanes run concurrently: one scratch buffer each)"""
        t = self._bwd_ws.get(lane)
        if t is None:
            t = self.arena.u8(int(lib.dpi_stats_workspace_bytes(self.max_C)))
            self._bwd_ws[lane] = t
        return t

    def map_tensor(self, layout: ChannelLayout) -> torch.Tensor:
        t = self._maps.get(layout)
        if t is None:
            t = torch.from_numpy(layout.phys2log()).to(self.device)
            self._maps[layout] = t
        return t

    def new_tensor(self, dims, layout: ChannelLayout, needs_grad: bool = True) -> Tn:
        s = Store(self, tuple(dims), layout.C_p)
        self.stores.append(s)
        return Tn(s, 0, layout, needs_grad)

    def register_grad_write(self, t: Tn, op: Op, tag: str):
        t.store.writers.append((t.coff, t.coff + t.C, op, tag))

    # ---- graph construction -------------------------------------------------------------------------
    def _unit(self, x: Tn, unit, out_layout: ChannelLayout, act, out: Optional[Tn] = None, feeds_conv: bool = False,
              lane: int = 0, emit_stats: bool = False) -> Tn:
        conv, bn = unit
        c = ConvOp(self, x, conv, out_layout, bn_follows=bn is not None)
        b = BnActOp(self, c.y, bn, act, out=out, round_out=feeds_conv, emit_stats=emit_stats)
        c.lane = b.lane = lane
        self.ops += [c, b]
        return b.out

    def _block(self, x: Tn, spec) -> Tn:
        act = self.net.spec["act"]
        c1, c2, c3 = (u[0].out_channels for u in (spec["conv3x3"], spec["conv5x5"], spec["conv7x7"]))
        parts = [ChannelLayout.dense(c1), ChannelLayout.dense(c2), ChannelLayout.dense(c3)]
        lay = ChannelLayout.concat(parts)
        offs = lay.part_offsets(parts)
        # The shortcut branch (1x1 conv + BN + act: HBM-bound) runs on lane 1, concurrently with the conv chain (tensor-
        # bound persistent kernels) on lane 0; fork where x is complete, join at the residual add.  Backward: the add
        # hands both branches their gradients (fork), the shortcut conv writes x.grad first and the chain's first conv
        # accumulates into it, so that dgrad waits for lane 1 (bwd_pre_wait).
        self.ops.append(MarkerOp(fwd=(1, 0)))
        # the three branch outputs stay in their own dense buffers (no concat buffer: see AddActOp)
        # (with a BatchNorm over the concatenation, each branch's BatchNorm + activation pass leaves the statistics of its
        #  output: AddActOp.part_stats)
        st = spec.get("bn1") is not None and _FUSE_PART_STATS
        o1 = self._unit(x, spec["conv3x3"], parts[0], act, feeds_conv=True, emit_stats=st)
        first_conv = self.ops[-2]
        o2 = self._unit(o1, spec["conv5x5"], parts[1], act, feeds_conv=True, emit_stats=st)
        o3 = self._unit(o2, spec["conv7x7"], parts[2], act, emit_stats=st)
        s = self._unit(x, spec["shortcut"], lay, act, lane=1)
        shortcut_bn = self.ops[-1]
        first_conv.bwd_pre_wait = (0, 1)
        first_conv.fuse_dgrad_with(self.ops[-2])        # conv3x3 + shortcut share x: one data-gradient launch
        self.ops.append(MarkerOp(fwd=(0, 1), bwd=(1, 0)))
        if spec.get("bn1") is not None:
            add = AddActOp(self, s, [o1, o2, o3], spec["bn1"], act, emit_stats=True, q_layout=lay)
            self.ops.append(add)
            fin = BnActOp(self, add.out, spec["bn2"], None, round_out=True)
            self.ops.append(fin)
            # backward: norm2's apply pass takes norm1's sums, norm1's apply pass the shortcut BatchNorm's
            fin.next_add = add
            add.next_bn = shortcut_bn
            return fin.out
        add = AddActOp(self, s, [o1, o2, o3], None, act, round_out=True, q_layout=lay)
        self.ops.append(add)
        return add.out

    def _respath(self, x: Tn, spec, out: Tn, lane: int = 0) -> Tn:
        act = self.net.spec["act"]
        f = spec["conv3x3"][0].out_channels
        lay = ChannelLayout.dense(f)
        # (the 3x3 unit is issued first: in the backward pass - reverse order - its conv then comes LAST, when the output
        #  gradients of both convs exist, and computes the data gradient of both in one launch)
        b = self._unit(x, spec["conv3x3"], lay, act, lane=lane)
        c3x3 = self.ops[-2]
        a = self._unit(x, spec["conv1x1"], lay, act, lane=lane)
        c3x3.fuse_dgrad_with(self.ops[-2])              # conv3x3 + conv1x1 share x: one data-gradient launch
        add = AddActOp(self, a, b, None, act, emit_stats=True)
        fin = BnActOp(self, add.out, spec["bn"], None, out=out, round_out=True)
        fin.next_add = add                              # backward: fin's apply pass writes both addends' gradients
        add.lane = fin.lane = lane
        self.ops += [add, fin]
        return fin.out

    def _level(self, x: Tn, levels, i: int) -> Tn:
        spec = levels[i]
        act = self.net.spec["act"]
        is3d = self.net.spec["is3d"]
        conv, bn = spec["down"]
        # The skip path (ResPath) of this level only meets the rest at the concat in front of the decoder block, so it
        # gets a lane of its own (2 + level) and runs concurrently with everything below this level - mostly small,
        # latency-bound launches that leave the machine half empty.  Backward: ResPath's convs write x.grad first, the
        # down conv accumulates into it and therefore waits for that lane.
        rlane = 2 + i
        self.ops.append(MarkerOp(fwd=(rlane, 0)))
        cdown = ConvOp(self, x, conv, x.layout, bn_follows=bn is not None)
        cdown.bwd_pre_wait = (0, rlane)
        self.ops.append(cdown)
        dact = BnActOp(self, cdown.y, bn, act, round_out=True)
        self.ops.append(dact)
        e = self._block(dact.out, spec["enc"])
        if i + 1 < len(levels):
            e = self._level(e, levels, i + 1)
        skip_lay = ChannelLayout.dense(spec["respath"]["conv3x3"][0].out_channels)
        cat_lay = ChannelLayout.concat([skip_lay, e.layout])
        cat = self.new_tensor(x.dims, cat_lay)
        self._respath(x, spec["respath"], cat.slice(0, skip_lay), lane=rlane)
        up = UpsampleOp(self, e, cat.slice(skip_lay.C_p, e.layout), self.net.spec["upsample"], up_d=is3d)
        for a in range(3):
            assert cat.dims[a] <= (2 * e.dims[a] if (is3d or a > 0) else e.dims[a]), "skip larger than upsampled branch"
        self.ops.append(up)
        self.ops.append(MarkerOp(fwd=(0, rlane), bwd=(rlane, 0)))
        return self._block(cat, spec["dec"])

    def _att_level(self, x: Tn, levels, i: int) -> Tn:
        """One scale of AttMulResUnet2D.forward (attention.py:249-262):
        ``x_dec = up_mb(concat([att(g, x), up(g)]))`` with ``g`` = the (decoded) tensor of the scale below."""
        spec = levels[i]
        act = self.net.spec["act"]
        conv, bn = spec["down"]
        cdown = ConvOp(self, x, conv, x.layout, bn_follows=True)                    # down<i>: conv s2 + BN + act
        dact = BnActOp(self, cdown.y, bn, act, round_out=True)
        self.ops += [cdown, dact]
        g = self._block(dact.out, spec["enc"])
        if i + 1 < len(levels):
            g = self._att_level(g, levels, i + 1)
        if any(x.dims[a] != 2 * g.dims[a] for a in (1, 2)):
            raise ValueError("attmultiunet: level sizes %s / %s — the up-sampled attention map must match the skip "
                             "tensor (attention.py:109-113); use spatial sizes divisible by %d" % (x.dims, g.dims, 2 ** len(levels)))
        # GridAttentionBlock.forward (attention.py:107-113)
        a = spec["att"]
        f_int = a["W_g"][0].out_channels
        lay = ChannelLayout.dense(f_int)
        g1 = self._unit(g, a["W_g"], lay, None)                                      # W_g: conv 1x1 + BN
        x1 = self._unit(x, a["W_x"], lay, None)                                      # W_x: conv 3x3 stride 2 + BN
        add = AddActOp(self, g1, x1, None, "ReLU", round_out=True)                   # relu(g1 + x1)
        one = ChannelLayout.dense(1)
        pconv = ConvOp(self, add.out, a["psi"], one, bn_follows=False)               # psi: conv 1x1 -> 1 channel
        sig = BnActOp(self, pconv.y, None, "Sigmoid")
        psi = self.new_tensor(x.dims, one)
        pup = UpsampleOp(self, sig.out, psi, "bilinear", up_d=False, feeds_conv=False)
        cat_lay = ChannelLayout.concat([x.layout, g.layout])
        cat = self.new_tensor(x.dims, cat_lay)                                       # concat([att(g, x), up(g)])
        gate = GateMulOp(self, x, psi, cat.slice(0, x.layout))
        up = UpsampleOp(self, g, cat.slice(x.layout.C_p, g.layout), self.net.spec["upsample"], up_d=False)
        self.ops += [add, pconv, sig, pup, gate, up]
        return self._block(cat, spec["dec"])

    def _build(self):
        spec = self.net.spec
        zl = ChannelLayout.dense(spec["inputdepth"])
        self.z = self.new_tensor(self.dims, zl, needs_grad=False)       # the fixed noise tensor z
        self.zin = self.new_tensor(self.dims, zl, needs_grad=False)     # z + sigma * eps  (main.py:148-150)
        x0 = self._block(self.zin, spec["first"])
        if spec.get("kind") == "attmultiunet":
            y0 = self._att_level(x0, spec["levels"], 0)
        else:
            y0 = self._level(x0, spec["levels"], 0)
        oc = spec["out"].out_channels
        ol = ChannelLayout.dense(oc)
        cout = ConvOp(self, y0, spec["out"], ol, bn_follows=False)
        self.ops.append(cout)
        self.out = cout.y
        if spec.get("last_act"):
            la = BnActOp(self, cout.y, None, spec["last_act"])
            self.ops.append(la)
            self.out = la.out
        self.out_layout = ol
        n = self.out.nvox * self.out.ld
        self.img = self.zeros(n)
        self.mask = self.zeros(n)
        self.best = self.zeros(n)
        self.loss_ws = torch.zeros(int(lib.dpi_loss_workspace_bytes()), dtype=torch.uint8, device=self.device)
        self.scalars = torch.zeros(8, dtype=torch.float64, device=self.device)
        self.hyper = torch.tensor([1e-3, 1.0], dtype=torch.float64, device=self.device)
        self.counter = torch.zeros(2, dtype=torch.int64, device=self.device)   # {iteration, noise seed}
        self.best_state = torch.zeros(2, dtype=torch.float64, device=self.device)
        self.history = torch.zeros((self.max_iters, 4), dtype=torch.float64, device=self.device)
        self.adam_m = torch.zeros_like(self.params.P)
        self.adam_v = torch.zeros_like(self.params.P)
        self.wgrad_ws = self.zeros(self.wgrad_ws_bytes // 4 + 4)
        self._bwd_ws: Dict[int, torch.Tensor] = {}
        # first-writer-overwrites / later-writers-accumulate, in BACKWARD order, alias-aware per Store
        for s in self.stores:
            written: List[Tuple[int, int]] = []
            for (lo, hi, op, tag) in reversed(s.writers):
                inside = [w for w in written if not (hi <= w[0] or lo >= w[1])]
                if inside:
                    cov = sorted(inside)
                    pos = lo
                    for (a, b) in cov:
                        if a > pos:
                            break
                        pos = max(pos, b)
                    assert pos >= hi, "partially initialised gradient slice"
                    op.acc[tag] = True
                else:
                    op.acc[tag] = False
                written.append((lo, hi))
        # every conv layer's weight pack (before the forward) and gradient un-pack (after the backward) is ONE launch
        # over a device-resident job table instead of 2 x 49 small launches
        convs = [op for op in self.ops if isinstance(op, ConvOp)]
        jobs = (_lib.PackJob * len(convs))(*[op.pack_job() for op in convs])
        self.pack_jobs = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(self.device)
        tf32 = 1 if self.prec == _lib.PREC_TF32 else 0
        self.pack_calls = [_Call("dpi_pack_conv_weights_batched", self.pack_jobs.data_ptr(), len(convs), tf32)]
        self.call_op: Dict[int, Op] = {}         # id(pre-marshalled call) -> the op it belongs to (profiling tools)

        def tagged(calls, op):
            for c in calls:
                if not isinstance(c, _Wait):
                    c.lane = op.lane
                    for cc in (c.calls if isinstance(c, _SideCall) else (c,)):
                        self.call_op[id(cc)] = op
            return calls

        self.fwd_calls = [c for op in self.ops for c in tagged(op.emit_fwd(), op)]
        self.bwd_calls = [c for op in reversed(self.ops) for c in tagged(op.emit_bwd(), op)]
        self.bwd_calls.append(_Call("dpi_unpack_conv_wgrad_batched", self.pack_jobs.data_ptr(), len(convs)))
        self.set_loss("mae")
        self.launches_per_iteration = None

    def set_loss(self, kind: str):
        """--loss mae|mse (parameter.py:82; main.py:24-27)"""
        if getattr(self, "loss_call", None) is not None and self.loss_kind == _lib.LOSS_CODES[kind]:
            return
        self.loss_kind = _lib.LOSS_CODES[kind]
        self.loss_call = self._make_loss_call(self.img, self.mask, self.scalars)
        for r in getattr(self, "rows", []):
            r.loss_call = self._make_loss_call(r.img, r.mask, r.scalars)
        self.graph = None

    def _make_loss_call(self, img: torch.Tensor, mask: torch.Tensor, scalars: torch.Tensor) -> _Call:
        nout = self.out.nvox * self.out.ld
        return _Call("dpi_masked_loss", self.out.ptr, img.data_ptr(), mask.data_ptr(), nout,
                     self.out.nvox * self.out_layout.C_l,
                     self.loss_kind | (_lib.ROUND_TF32 if self.prec == _lib.PREC_TF32 else 0), self.out.gptr,
                     self.loss_ws.data_ptr(), self.loss_ws.numel(), scalars.data_ptr())

    # ---- data movement ------------------------------------------------------------------------------------
    @property
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _to_cl(self, src: torch.Tensor, dst_ptr: int, layout: ChannelLayout, ld: int, st=None):
        src = src.to(device=self.device, dtype=torch.float32).contiguous()
        nvox = self.dims[0] * self.dims[1] * self.dims[2]
        assert src.numel() == layout.C_l * nvox, "shape mismatch: %s vs C=%d dims=%s" % (tuple(src.shape), layout.C_l, self.dims)
        _lib.call("dpi_nchw_to_cl", _vp(src.data_ptr()), layout.C_l, nvox, _vp(self.map_tensor(layout).data_ptr()),
                  _vp(dst_ptr), ld, layout.C_p, _vp(st if st is not None else self.stream))
        self._last_src = src   # keep alive until the stream has consumed it

    def _from_cl(self, src_ptr: int, layout: ChannelLayout, ld: int, shape) -> torch.Tensor:
        nvox = self.dims[0] * self.dims[1] * self.dims[2]
        dst = torch.empty((layout.C_l, nvox), dtype=torch.float32, device=self.device)
        _lib.call("dpi_cl_to_nchw", _vp(src_ptr), ld, layout.C_p, _vp(self.map_tensor(layout).data_ptr()),
                  _vp(dst.data_ptr()), layout.C_l, nvox, _vp(self.stream))
        return dst.view(shape)

    def set_noise_input(self, z_nchw: torch.Tensor):
        """load the fixed input noise z (main.py:59-64), NCDHW / NCHW"""
        self._to_cl(z_nchw, self.z.ptr, self.z.layout, self.z.ld)

    def set_data_forgetting(self, add_data_nchw: Optional[torch.Tensor], weights=None):
        """``--data_forgetting_factor`` (main.py:86-97,153-155): ``add_data_nchw`` holds the image channels of the
        normalised decimated data (its repetition along the input depth is done by the kernel), ``weights`` the per-
        iteration factors; ``None`` switches the option off.  Changes the launch list, so the graph is re-captured."""
        if add_data_nchw is None:
            if getattr(self, "_forget", None) is not None:
                self._forget = None
                self.graph = None
            return
        w = torch.as_tensor(np.asarray(weights, dtype=np.float64)).to(torch.float32)    # the reference multiplies in fp32
        f = getattr(self, "_forget", None)
        if f is None or f[1].numel() != w.numel():
            # the captured graph holds these two pointers: they stay for the life of the engine
            f = (self.zeros(self.out.nvox * self.out.ld), torch.zeros(w.numel(), dtype=torch.float32, device=self.device))
            self._forget = f
            self.graph = None
        self._to_cl(add_data_nchw, f[0].data_ptr(), self.out_layout, self.out.ld)
        f[1].copy_(w)

    def add_forgetting_data(self, st=None):
        f = getattr(self, "_forget", None)
        if f is None:
            return
        _lib.call("dpi_add_data_dev", _vp(self.zin.ptr), self.zin.ld, self.zin.layout.C_l, self.zin.nvox,
                  _vp(f[0].data_ptr()), self.out.ld, self.out_layout.C_l, _vp(f[1].data_ptr()), int(f[1].numel()),
                  _vp(self.counter.data_ptr()), 1 if self.prec == _lib.PREC_TF32 else 0,
                  _vp(self.stream if st is None else st))

    def network_input_nchw(self) -> torch.Tensor:
        """the perturbed input the last forward pass consumed (the entries of ``input_list``, main.py:155)"""
        shape = (1, self.zin.layout.C_l) + (self.dims if self.net.spec["is3d"] else self.dims[1:])
        return self._from_cl(self.zin.ptr, self.zin.layout, self.zin.ld, shape)

    def set_network_input(self, x_nchw: torch.Tensor):
        self._to_cl(x_nchw, self.zin.ptr, self.zin.layout, self.zin.ld)

    def set_target(self, img_nchw: torch.Tensor, mask_nchw: torch.Tensor):
        """img_ and mask_ of main.py:134-135"""
        self._to_cl(img_nchw, self.img.data_ptr(), self.out_layout, self.out.ld)
        self._to_cl(mask_nchw, self.mask.data_ptr(), self.out_layout, self.out.ld)

    def output_nchw(self, best: bool = False) -> torch.Tensor:
        shape = (1, self.out_layout.C_l) + (self.dims if self.net.spec["is3d"] else self.dims[1:])
        return self._from_cl(self.best.data_ptr() if best else self.out.ptr, self.out_layout, self.out.ld, shape)

    # ---- execution ------------------------------------------------------------------------------------------
    def _refresh(self):
        if self.params.stale():
            self.params.adopt()

    def _run(self, calls, st=None):
        """Issue a launch list: lane 0 on the current stream, lane 1 (shortcut branches) and the weight-gradient lane on
        the engine's side streams, ordered by the _Wait entries; everything is joined back into lane 0 at the end.
        With a foreign raw stream handle, or DPI_SIDE_STREAM=0, the list is issued in order on one stream (the list
        order is a valid serial schedule)."""
        main = torch.cuda.current_stream(self.device)
        if self.side_streams is None or (st is not None and int(st) != main.cuda_stream):
            stp = _vp(main.cuda_stream if st is None else st)
            for c in calls:
                if isinstance(c, _Wait):
                    continue
                for cc in (c.calls if isinstance(c, _SideCall) else (c,)):
                    cc(stp)
            return
        streams = [main] + self.side_streams            # [lane 0, lanes 1.., weight-gradient lane]
        ptrs = [_vp(t.cuda_stream) for t in streams]
        wg = len(streams) - 1
        dirty = set()

        def order(waiter: int, signaler: int):
            if waiter == signaler:
                return
            ev = torch.cuda.Event()
            ev.record(streams[signaler])
            streams[waiter].wait_event(ev)

        for c in calls:
            if isinstance(c, _Wait):
                order(c.waiter, c.signaler)
                dirty.add(c.waiter)
            elif isinstance(c, _SideCall):
                order(wg, c.lane)                        # the operands of this weight gradient are final here
                dirty.add(wg)
                for cc in c.calls:
                    cc(ptrs[wg])
            else:
                c(ptrs[c.lane])
                if c.lane:
                    dirty.add(c.lane)
        for lane in sorted(dirty):
            order(0, lane)

    def run_forward(self, st=None):
        stp = _vp(self.stream if st is None else st)
        for c in self.pack_calls:
            c(stp)
        self._run(self.fwd_calls, st)

    def run_loss(self, st=None):
        self.loss_call(_vp(self.stream if st is None else st))

    def run_backward(self, st=None):
        self._run(self.bwd_calls[:-1], st)
        self.bwd_calls[-1](_vp(self.stream if st is None else st))   # gradient un-pack, after the join

    def perturb_input(self, sigma: float, eps_nchw: Optional[torch.Tensor] = None, seed: int = 0, st=None):
        """input_ = z + reg_noise_std * N(0,1)  (main.py:148-150); eps supplied for parity runs"""
        stp = _vp(self.stream if st is None else st)
        n = self.z.nvox * self.z.ld
        if eps_nchw is not None:
            if not hasattr(self, "_eps"):
                self._eps = self.zeros(n)
            self._to_cl(eps_nchw, self._eps.data_ptr(), self.z.layout, self.z.ld)
            _lib.call("dpi_noise_axpy", _vp(self.z.ptr), _vp(self._eps.data_ptr()), _vp(self.zin.ptr), n,
                      float(sigma), 0, 0, 1 if self.prec == _lib.PREC_TF32 else 0, stp)
        else:
            _lib.call("dpi_noise_axpy_dev", _vp(self.z.ptr), _vp(self.zin.ptr), n, float(sigma), int(seed),
                      _vp(self.counter.data_ptr()), 1 if self.prec == _lib.PREC_TF32 else 0, stp)

    def adam_step(self, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, st=None):
        P = self.params
        _lib.call("dpi_adam_step_dev", _vp(P.P.data_ptr()), _vp(P.G.data_ptr()), _vp(self.adam_m.data_ptr()),
                  _vp(self.adam_v.data_ptr()), P.n, _vp(self.hyper.data_ptr()), float(betas[0]), float(betas[1]),
                  float(eps), float(weight_decay), _vp(self.stream if st is None else st))

    def iteration_end(self, st=None):
        n = self.out.nvox * self.out.ld
        _lib.call("dpi_iteration_end", _vp(self.scalars.data_ptr()), _vp(self.hyper.data_ptr()),
                  _vp(self.counter.data_ptr()), _vp(self.history.data_ptr()), self.max_iters,
                  _vp(self.best_state.data_ptr()), _vp(self.out.ptr), _vp(self.best.data_ptr()), n,
                  _vp(self.stream if st is None else st))

    # ---- batch rows (shared-network mode) ----------------------------------------------------------------------
    def new_row(self, seed: int = 0) -> Row:
        if not hasattr(self, "rows"):
            self.rows: List[Row] = []
        r = Row(self, primary=not self.rows, seed=seed)
        self.rows.append(r)
        return r

    def row_load(self, row: Row, z_nchw: torch.Tensor, img_nchw: torch.Tensor, mask_nchw: torch.Tensor):
        self._to_cl(z_nchw, row.z.data_ptr(), self.z.layout, self.z.ld)
        self._to_cl(img_nchw, row.img.data_ptr(), self.out_layout, self.out.ld)
        self._to_cl(mask_nchw, row.mask.data_ptr(), self.out_layout, self.out.ld)

    def run_row(self, row: Row, sigma: float, st=None):
        """perturb this row's z, forward, masked loss (+ metrics) of this row, backward: leaves the row's gradient in
        the flat gradient buffer and its {loss, snr, pcorr} in ``row.scalars``"""
        stp = _vp(self.stream if st is None else st)
        n = self.z.nvox * self.z.ld
        tf32 = 1 if self.prec == _lib.PREC_TF32 else 0
        if sigma > 0:
            _lib.call("dpi_noise_axpy_dev", _vp(row.z.data_ptr()), _vp(self.zin.ptr), n, float(sigma), 0,
                      _vp(row.counter.data_ptr()), tf32, stp)
        else:
            _lib.call("dpi_copy_slice", _vp(row.z.data_ptr()), self.z.ld, _vp(self.zin.ptr), self.zin.ld, self.z.nvox,
                      self.z.C, 0, stp)
        self.run_forward(st)
        row.loss_call(stp)
        self.run_backward(st)

    def row_end(self, row: Row, st=None):
        """bookkeeping of one row: history row, best-output tracking, noise counter + 1 (and, for row 0, Adam step + 1)"""
        n = self.out.nvox * self.out.ld
        _lib.call("dpi_iteration_end", _vp(row.scalars.data_ptr()), _vp(row.hyper.data_ptr()),
                  _vp(row.counter.data_ptr()), _vp(row.history.data_ptr()), self.max_iters,
                  _vp(row.best_state.data_ptr()), _vp(self.out.ptr), _vp(row.best.data_ptr()), n,
                  _vp(self.stream if st is None else st))

    def row_output_nchw(self, row: Row) -> torch.Tensor:
        shape = (1, self.out_layout.C_l) + (self.dims if self.net.spec["is3d"] else self.dims[1:])
        return self._from_cl(row.best.data_ptr(), self.out_layout, self.out.ld, shape)

    def reset_loop_state(self, lr: float, seed: int = 0):
        self.hyper.copy_(torch.tensor([lr, 1.0], dtype=torch.float64))
        self.counter.copy_(torch.tensor([0, int(seed)], dtype=torch.int64))
        self.best_state.zero_()
        self.history.zero_()
        self.adam_m.zero_()
        self.adam_v.zero_()

    def set_lr(self, lr: float):
        self.hyper[0:1].copy_(torch.tensor([lr], dtype=torch.float64), non_blocking=True)

    def iteration(self, sigma: float, seed: int = 0, st=None):
        """One full optimisation iteration (main.py:210-213): perturb, forward, loss, backward, Adam, bookkeeping."""
        if sigma > 0:
            self.perturb_input(sigma, None, seed, st)
        else:
            _lib.call("dpi_copy_slice", _vp(self.z.ptr), self.z.ld, _vp(self.zin.ptr), self.zin.ld, self.z.nvox,
                      self.z.C, 0, _vp(self.stream if st is None else st))
        self.add_forgetting_data(st)
        self.run_forward(st)
        self.run_loss(st)
        self.run_backward(st)
        self.adam_step(st=st)
        self.iteration_end(st)

    def capture(self, sigma: float, seed: int = 0):
        """Capture one iteration into a CUDA graph (all launch arguments are fixed device pointers)."""
        self._refresh()
        # (no warm-up iteration is needed: every launch argument is a fixed device pointer or scalar, the kernels hold no
        #  lazily-created device state, and the one-time cudaFuncSetAttribute calls are not stream operations)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            n0 = int(lib.dpi_launch_count())
            with torch.cuda.graph(g, stream=s):
                self.iteration(sigma, seed, st=torch.cuda.current_stream(self.device).cuda_stream)
            self.launches_per_iteration = int(lib.dpi_launch_count()) - n0
        torch.cuda.current_stream(self.device).wait_stream(s)
        self.graph = g
        self._graph_sigma = float(sigma)
        return g

    def read_scalars(self) -> Tuple[float, float, float]:
        v = self.scalars.cpu()
        return float(v[0]), float(v[1]), float(v[2])
