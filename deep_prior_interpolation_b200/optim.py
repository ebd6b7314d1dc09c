"""Fused flat Adam as a ``torch.optim.Optimizer`` (replaces ``torch.optim.Adam`` of ``main.py:200``).

``param_groups[0]['lr']`` is read (``main.py:169``) and written (``ReduceLROnPlateau``, ``main.py:201-204``) by the
reference's loop, so this stays a real ``Optimizer`` subclass; ``step()`` is ONE kernel launch over the engine's
flat parameter / gradient / moment buffers (``dpi_adam_step``), with the exact update order of
``torch.optim.Adam`` (single-tensor path).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, net, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        from .architectures import DeepPriorNet
        if not isinstance(net, DeepPriorNet):
            raise TypeError("FusedAdam optimises a DeepPriorNet (pass the network, not net.parameters())")
        self.net = net
        super().__init__(list(net.parameters()), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._t = 0
        self._m = self._v = None

    def zero_grad(self, set_to_none: bool = True):
        # the engine overwrites the flat gradient buffer every backward; nothing to clear
        for p in self.param_groups[0]["params"]:
            if p.grad is not None and set_to_none:
                p.grad = None

    @torch.no_grad()
    def step(self, closure=None):
        eng = self.net._engine
        if eng is None:
            raise RuntimeError("FusedAdam.step() before any forward/backward")
        P = eng.params
        grp = self.param_groups[0]
        # gradients normally alias the flat buffer; gather any that do not (e.g. user-assigned .grad)
        for p, off in zip(P.plist, P.poff):
            if p.grad is not None and p.grad.data_ptr() != P.G.data_ptr() + 4 * off:
                P.G[off:off + p.numel()].view(p.shape).copy_(p.grad)
        if self._m is None or self._m.numel() != P.P.numel() or self._m.device != P.P.device:
            self._m, self._v = torch.zeros_like(P.P), torch.zeros_like(P.P)
        self._t += 1
        vp = lambda t: C.c_void_p(t.data_ptr())
        _lib.call("dpi_adam_step", vp(P.P), vp(P.G), vp(self._m), vp(self._v), P.n, float(grp["lr"]),
                  float(grp["betas"][0]), float(grp["betas"][1]), float(grp["eps"]), float(grp["weight_decay"]), self._t,
                  C.c_void_p(torch.cuda.current_stream(P.P.device).cuda_stream))
        return None
