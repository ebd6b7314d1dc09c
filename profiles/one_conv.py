#!/usr/bin/env python
"""One convolution launch at a chosen shape, timed with CUDA events (and a convenient target for `ncu -k regex:...`):

    python profiles/one_conv.py --op dgrad_fused --dims 256 128 128 --cin 72 --cout 4 --cout_b 28 [--reps 5]
    python profiles/one_conv.py --op fwd|dgrad|wgrad --dims ... --cin C --cout N [--k 3] [--stride 1] [--acc 1]

Channel counts are the padded (multiple of 4) ones the engine uses."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--op", default="fwd")
    ap.add_argument("--dims", type=int, nargs=3, default=[256, 128, 128])
    ap.add_argument("--cin", type=int, default=28)
    ap.add_argument("--cout", type=int, default=16)
    ap.add_argument("--cout_b", type=int, default=0)
    ap.add_argument("--k", type=int, default=3)
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--acc", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from deep_prior_interpolation_b200 import _lib
    dev = torch.device("cuda", 0)
    D, H, W = a.dims
    k = a.k
    taps = k ** 3
    cc = a.cout + a.cout_b
    geom = _lib.ConvGeom(D, H, W, a.cin, cc, k, k, k, a.stride)
    od = [(n + 2 * (k // 2) - k) // a.stride + 1 for n in a.dims]
    nvi, nvo = D * H * W, od[0] * od[1] * od[2]
    x = torch.randn(nvi, a.cin, device=dev)
    y = torch.randn(nvo, cc, device=dev)
    w = torch.randn(cc * taps * a.cin, device=dev) * 0.05
    b = torch.zeros(cc, device=dev)
    vp = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ws = torch.zeros(int(_lib.lib.dpi_conv_wgrad_workspace_bytes(C.byref(geom))) // 4 + 4, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)

    def run():
        if a.op == "fwd":
            _lib.call("dpi_conv_fwd", vp(x), a.cin, vp(w), vp(b), vp(y), cc, C.byref(geom), 1, st)
        elif a.op == "dgrad":
            _lib.call("dpi_conv_dgrad", vp(y), cc, vp(w), vp(x), a.cin, C.byref(geom), a.acc, 1, st)
        elif a.op == "dgrad_fused":
            _lib.call("dpi_conv_dgrad_fused", vp(y), cc, vp(w), vp(x), a.cin, C.byref(geom), a.cout, a.acc, 1, st)
        else:
            _lib.call("dpi_conv_wgrad", vp(x), a.cin, vp(y), cc, vp(w), C.byref(geom), vp(ws), ws.numel() * 4, 1, st)

    ts = []
    for i in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("%s dims %s cin %d cout %d+%d k%d s%d acc %d: %s us (min %.1f)" % (a.op, a.dims, a.cin, a.cout, a.cout_b, k, a.stride,
                                                                            a.acc, " ".join("%.1f" % t for t in ts), min(ts)))


if __name__ == "__main__":
    main()
