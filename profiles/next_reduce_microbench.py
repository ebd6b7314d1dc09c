#!/usr/bin/env python
"""Time the BatchNorm-backward apply pass with and without the next unit's reduce fused in (dpi_bn_next_reduce), at the
full-resolution shapes of a MultiRes block: C = 28 (three parts 4/8/16), 256x128x128 voxels.  L2 flushed between reps.
DPI_B200_LIB selects a library variant (compile-time unroll / launch bounds of the fused kernels)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from deep_prior_interpolation_b200 import _lib
    dev = torch.device("cuda", 0)
    nvox, widths = 256 * 128 * 128, [4, 8, 16]
    Cc = sum(widths)
    vp = lambda t: C.c_void_p(t.data_ptr() if t is not None else None)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rnd = lambda *sh: torch.randn(*sh, device=dev)
    dy, t, y, xs = rnd(nvox, Cc), rnd(nvox, Cc), rnd(nvox, Cc), rnd(nvox, Cc)
    qs = [rnd(nvox, w) for w in widths]
    dqs = [torch.zeros(nvox, w, device=dev) for w in widths]
    parts = _lib.Parts.make([q.data_ptr() for q in qs], widths, widths)
    dparts = _lib.Parts.make([q.data_ptr() for q in dqs], widths, widths)
    xparts = _lib.Parts.make([xs.data_ptr()], [Cc], [Cc])
    mean, invstd, scale, shift, c1, c2 = (rnd(Cc) for _ in range(6))
    dx, dp = torch.zeros(nvox, Cc, device=dev), torch.zeros(nvox, Cc, device=dev)
    ws = torch.zeros(int(_lib.lib.dpi_stats_workspace_bytes(Cc)), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    nx1 = _lib.NextReduce.make(1, 1, parts, mean.data_ptr(), invstd.data_ptr(), 0, 0, ws.data_ptr())
    nx2 = _lib.NextReduce.make(2, 1, xparts, mean.data_ptr(), invstd.data_ptr(), scale.data_ptr(), shift.data_ptr(), ws.data_ptr())

    cases = {
        "apply (norm2)": lambda: _lib.call("dpi_bn_bwd_apply", vp(dy), Cc, None, Cc, 0, vp(t), Cc, vp(mean), vp(invstd), vp(scale),
                                           None, vp(c1), vp(c2), vp(dx), Cc, nvox, Cc, 0, st),
        "reduce_parts (norm1)": lambda: _lib.call("dpi_bn_bwd_reduce_parts", vp(dx), Cc, vp(t), Cc, 1, parts, vp(mean), vp(invstd),
                                                  nvox, Cc, vp(ws), st),
        "apply_next kind 1": lambda: _lib.call("dpi_bn_bwd_apply_next", vp(dy), Cc, None, Cc, 0, vp(t), Cc, vp(mean), vp(invstd),
                                               vp(scale), None, vp(c1), vp(c2), vp(dx), Cc, nvox, Cc, 0, nx1, st),
        "apply_parts + dp (norm1)": lambda: _lib.call("dpi_bn_bwd_apply_parts", vp(dy), Cc, vp(y), Cc, 1, parts, vp(mean), vp(invstd),
                                                      vp(scale), vp(c1), vp(c2), dparts, 0, vp(dp), Cc, nvox, Cc, st),
        "reduce (shortcut BN)": lambda: _lib.call("dpi_bn_bwd_reduce", vp(dp), Cc, None, Cc, 1, vp(xs), Cc, vp(mean), vp(invstd),
                                                  vp(scale), vp(shift), nvox, Cc, vp(ws), st),
        "apply_parts_next kind 2": lambda: _lib.call("dpi_bn_bwd_apply_parts_next", vp(dy), Cc, vp(y), Cc, 1, parts, vp(mean),
                                                     vp(invstd), vp(scale), vp(c1), vp(c2), dparts, 0, vp(dp), Cc, nvox, Cc, nx2, st),
    }
    print("library: %s" % _lib.LIB_PATH)
    for name, fn in cases.items():
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        print("  %-28s %7.1f us" % (name, min(ts[1:])))


if __name__ == "__main__":
    main()
