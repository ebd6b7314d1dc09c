"""Per-kernel parity tests through the C ABI (ctypes), against plain PyTorch references of the same op.

Tolerances: exact-fp32 kernels are compared with fp64 references at rtol 2e-5 of the tensor scale; the
tcgen05 TF32 path at 2e-3 (TF32 has a 10-bit mantissa).
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _imports():
    from deep_prior_interpolation_b200 import _lib
    from deep_prior_interpolation_b200.layout import ChannelLayout, pad4
    return _lib, ChannelLayout, pad4


def vp(t):
    return C.c_void_p(t.data_ptr() if t is not None else None)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_cl(x, Cp):
    """(C, D, H, W) -> [nvox, Cp] channels-last with zero pads"""
    c = x.shape[0]
    out = torch.zeros((x[0].numel(), Cp), dtype=torch.float32, device=x.device)
    out[:, :c] = x.reshape(c, -1).t()
    return out.contiguous()


def from_cl(t, c, dims):
    return t[:, :c].t().reshape((c,) + tuple(dims))


def pack_w(w, Cout_p, Cin_p):
    """[Cout, Cin, *k] -> [Cout_p][taps][Cin_p], and the dgrad pack [Cin_p][taps][Cout_p]"""
    co, ci = w.shape[:2]
    taps = int(np.prod(w.shape[2:]))
    wf = torch.zeros((Cout_p, taps, Cin_p), dtype=torch.float32, device=w.device)
    wf[:co, :, :ci] = w.reshape(co, ci, taps).permute(0, 2, 1)
    wd = wf.permute(2, 1, 0).contiguous()
    return wf.contiguous(), wd


CONV_CASES = [
    # (D,H,W, Cin, Cout, k(3d tuple), stride)
    ((6, 9, 10), 8, 4, (3, 3, 3), 1),
    ((5, 7, 9), 25, 16, (3, 3, 3), 1),
    ((8, 8, 8), 64, 25, (1, 1, 1), 1),
    ((7, 10, 9), 13, 13, (3, 3, 3), 2),
    ((8, 6, 12), 67, 4, (3, 3, 3), 1),
    ((1, 17, 13), 17, 26, (1, 3, 3), 1),
    ((1, 17, 13), 25, 25, (1, 3, 3), 2),
    ((4, 4, 4), 142, 71, (3, 3, 3), 1),
    # shapes of the grid-attention blocks (attention.py:86-105): W_x = 3x3 stride 2 with Cin != Cout, W_g = wide 1x1,
    # psi = 1x1 to one channel
    ((1, 16, 12), 51, 64, (1, 3, 3), 2),
    ((1, 8, 6), 212, 256, (1, 3, 3), 2),
    ((1, 6, 4), 426, 256, (1, 1, 1), 1),
    ((1, 24, 16), 64, 1, (1, 1, 1), 1),
]


def _conv_ref(x, w, b, k, stride):
    pad = tuple((kk - 1) // 2 for kk in k)
    st = tuple(stride if kk > 1 else 1 for kk in k)
    return F.conv3d(x[None], w, b, stride=st, padding=pad)[0]


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(case, prec):
    _run_conv_case(case, prec)


# shapes sized so that the persistent column-march kernel (conv_tc_march.cu) runs several work units per CTA,
# segments longer than one plane, partial edge tiles, the 2-D (kd = 1) form and the wide-N dgrad (N = 72 -> BN 80)
MARCH_CASES = [
    ((20, 40, 24), 8, 13, (3, 3, 3), 1),
    ((12, 16, 8), 72, 4, (3, 3, 3), 1),
    ((9, 33, 17), 25, 16, (3, 3, 3), 1),
    ((1, 40, 30), 17, 26, (1, 3, 3), 1),
    ((16, 16, 16), 4, 8, (3, 3, 3), 1),
    ((10, 20, 12), 40, 36, (3, 3, 3), 1),     # two channel blocks x two output-channel blocks
    ((12, 20, 18), 25, 25, (3, 3, 3), 2),     # stride-2 dgrad: one march per output parity class
    ((11, 19, 17), 51, 51, (3, 3, 3), 2),     # odd sizes, two channel chunks
    ((1, 40, 30), 25, 25, (1, 3, 3), 2),      # 2-D stride 2
    ((6, 20, 12), 67, 25, (1, 1, 1), 1),      # 1x1x1 through the march pipeline (bare tiles), three channel chunks
    ((1, 33, 17), 51, 32, (1, 1, 1), 1),      # 2-D 1x1, partial tiles
    ((8, 16, 16), 137, 8, (3, 3, 3), 1),      # dgrad with N = 140 > 128: two launches over output-channel ranges (72 + 68)
    ((6, 16, 8), 137, 51, (1, 1, 1), 1),      # the same for a 1x1 (decoder shortcut of level 1), TMA-store epilogue
    ((4, 16, 8), 51, 276, (1, 1, 1), 1),      # forward with N = 276: three ranges of 92
    # four output channels -> role-swapped kernel (conv_tc_swap.cu): several tiles / segments / channel chunks, partial
    # edge tiles, one logical output channel (the final conv), and through the dgrad of 4 -> N convs (C = 8 / 16 / 28)
    ((20, 40, 24), 64, 4, (3, 3, 3), 1),
    ((9, 33, 17), 25, 1, (3, 3, 3), 1),
    ((7, 18, 11), 4, 13, (3, 3, 3), 1),
    ((5, 16, 8), 4, 25, (3, 3, 3), 1),
    # >= 128 K voxels: the reduction of a wide-input, narrow-output conv whose weights do not fit is split in two channel
    # ranges (forward), and a narrow output that only fits with 32-byte rows in two output-channel ranges (51 -> 32)
    ((32, 64, 64), 137, 8, (3, 3, 3), 1),
    ((32, 64, 64), 51, 32, (3, 3, 3), 1),
]


@pytest.mark.parametrize("ctas", ["3", "148"])
@pytest.mark.parametrize("case", MARCH_CASES)
def test_conv_march_persistent(case, ctas, monkeypatch):
    monkeypatch.setenv("DPI_TC_MARCH_CTAS", ctas)
    monkeypatch.setenv("DPI_TC_WGRAD_MARCH_CTAS", ctas)   # same for the marching weight-gradient kernel
    _run_conv_case(case, 1)


@pytest.mark.parametrize("case", [c for c in CONV_CASES + MARCH_CASES if c[4] == 2])
def test_stride2_convs_on_the_other_kernels(case, monkeypatch):
    """the stride-2 convs have two tcgen05 paths each: forward - four parity-class boxes per plane on the halo kernel
    (default) or one box per tap (conv_tc_kernel); data gradient - one launch with eight accumulators (default up to 32 K
    u positions) or one march launch per output parity class.  The default paths run in the tests above; here the others."""
    monkeypatch.setenv("DPI_TC_HALO_S2", "0")
    monkeypatch.setenv("DPI_TC_HALO_S2T", "0")
    _run_conv_case(case, 1)


def _run_conv_case(case, prec):
    _lib, ChannelLayout, pad4 = _imports()
    dims, cin, cout, k, stride = case
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = torch.randn((cin,) + dims, generator=g, dtype=torch.float64)
    w = torch.randn((cout, cin) + k, generator=g, dtype=torch.float64) * 0.1
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    x.requires_grad_(True)
    w.requires_grad_(True)
    y = _conv_ref(x, w, b, k, stride)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    Cip, Cop = pad4(cin), pad4(cout)
    xcl = to_cl(x.detach().float().to(dev), Cip)
    wf, wd = pack_w(w.detach().float().to(dev), Cop, Cip)
    bp = torch.zeros(Cop, device=dev)
    bp[:cout] = b.float().to(dev)
    odims = tuple(y.shape[1:])
    ycl = torch.full((int(np.prod(odims)), Cop), 7.0, device=dev)
    geom = _lib.ConvGeom(dims[0], dims[1], dims[2], Cip, Cop, k[0], k[1], k[2], stride)
    tol = 2e-5 if prec == 0 else 3e-3

    _lib.call("dpi_conv_fwd", vp(xcl), Cip, vp(wf), vp(bp), vp(ycl), Cop, C.byref(geom), prec, stream())
    got = from_cl(ycl, cout, odims).double().cpu()
    scale = y.detach().abs().max().item()
    assert (got - y.detach()).abs().max().item() <= tol * scale, "forward"
    if Cop > cout:
        assert ycl[:, cout:].abs().max().item() == 0.0, "pad channels must stay zero"

    # the same conv leaving the BatchNorm partial sums of y in a stats workspace (fused into the tcgen05 march
    # epilogue for Cout <= 32, a separate pass otherwise): identical y, and sums equal to fp64 sums of the stored y
    ycl2 = torch.full_like(ycl, -3.0)
    sws = torch.zeros(int(_lib.lib.dpi_stats_workspace_bytes(Cop)), dtype=torch.uint8, device=dev)
    _lib.call("dpi_conv_fwd_stats", vp(xcl), Cip, vp(wf), vp(bp), vp(ycl2), Cop, C.byref(geom), prec, vp(sws), stream())
    assert torch.equal(ycl, ycl2), "forward with fused statistics must store the same values"
    hdr = sws[:16].view(torch.int64).cpu()
    nrows, cc = int(hdr[0]), int(hdr[1])
    assert cc == Cop and 1 <= nrows <= 592
    rows = sws[16:16 + nrows * 2 * Cop * 8].view(torch.float64).view(nrows, 2, Cop).sum(0).cpu()
    y64 = ycl.double()
    s_ref, q_ref = y64.sum(0).cpu(), (y64 * y64).sum(0).cpu()
    assert (rows[0] - s_ref).abs().max().item() <= 1e-11 * (y64.abs().sum(0).max().item() + 1), "sum(y)"
    assert (rows[1] - q_ref).abs().max().item() <= 1e-11 * (q_ref.max().item() + 1), "sum(y^2)"

    dycl = to_cl(dy.float().to(dev), Cop)
    dxcl = torch.full_like(xcl, 3.0)
    _lib.call("dpi_conv_dgrad", vp(dycl), Cop, vp(wd), vp(dxcl), Cip, C.byref(geom), 0, prec, stream())
    gotdx = from_cl(dxcl, cin, dims).double().cpu()
    sdx = x.grad.abs().max().item()
    assert (gotdx - x.grad).abs().max().item() <= tol * sdx, "dgrad"
    # accumulate flag
    _lib.call("dpi_conv_dgrad", vp(dycl), Cop, vp(wd), vp(dxcl), Cip, C.byref(geom), 1, prec, stream())
    gotdx2 = from_cl(dxcl, cin, dims).double().cpu()
    assert (gotdx2 - 2 * x.grad).abs().max().item() <= 2 * tol * sdx, "dgrad accumulate"

    ws_bytes = int(_lib.lib.dpi_conv_wgrad_workspace_bytes(C.byref(geom)))
    ws = torch.zeros(ws_bytes // 4 + 4, device=dev)
    dwp = torch.zeros_like(wf)
    _lib.call("dpi_conv_wgrad", vp(xcl), Cip, vp(dycl), Cop, vp(dwp), C.byref(geom), vp(ws), ws.numel() * 4, prec,
              stream())
    taps = int(np.prod(k))
    gotdw = dwp[:cout, :, :cin].permute(0, 2, 1).reshape(cout, cin, taps).double().cpu()
    refdw = w.grad.reshape(cout, cin, taps)
    assert (gotdw - refdw).abs().max().item() <= tol * refdw.abs().max().item(), "wgrad"
    # bit-reproducible
    dwp2 = torch.zeros_like(wf)
    _lib.call("dpi_conv_wgrad", vp(xcl), Cip, vp(dycl), Cop, vp(dwp2), C.byref(geom), vp(ws), ws.numel() * 4, prec,
              stream())
    assert torch.equal(dwp, dwp2), "wgrad must be deterministic"


def test_pack_unpack_weights():
    _lib, ChannelLayout, pad4 = _imports()
    dev = torch.device("cuda")
    lay_in = ChannelLayout.concat([ChannelLayout.dense(4), ChannelLayout.dense(8), ChannelLayout.dense(13)])
    lay_out = ChannelLayout.dense(17)
    cin, cout, taps = lay_in.C_l, lay_out.C_l, 27
    w = torch.randn(cout, cin, 3, 3, 3, device=dev)
    b = torch.randn(cout, device=dev)
    mi = torch.from_numpy(lay_in.phys2log()).to(dev)
    mo = torch.from_numpy(lay_out.phys2log()).to(dev)
    wf = torch.full((lay_out.C_p, taps, lay_in.C_p), 9.0, device=dev)
    wd = torch.full((lay_in.C_p, taps, lay_out.C_p), 9.0, device=dev)
    bp = torch.full((lay_out.C_p,), 9.0, device=dev)
    _lib.call("dpi_pack_conv_weights", vp(w), vp(mo), vp(mi), cout, cin, lay_out.C_p, lay_in.C_p, taps, vp(wf), vp(wd),
              vp(b), vp(bp), 0, stream())
    ref = torch.zeros_like(wf)
    wl = w.reshape(cout, cin, taps)
    for p_o in range(lay_out.C_p):
        lo = int(mo[p_o])
        for p_i in range(lay_in.C_p):
            li = int(mi[p_i])
            if lo >= 0 and li >= 0:
                ref[p_o, :, p_i] = wl[lo, li]
    assert torch.equal(wf, ref)
    assert torch.equal(wd, ref.permute(2, 1, 0).contiguous())
    assert torch.equal(bp[:cout], b) and bp[cout:].abs().max().item() == 0
    dw = torch.zeros_like(w)
    _lib.call("dpi_unpack_conv_wgrad", vp(wf), vp(mo), vp(mi), cout, cin, lay_out.C_p, lay_in.C_p, taps, vp(dw), stream())
    assert torch.equal(dw, w)


@pytest.mark.parametrize("C_l,nvox", [(4, 1000), (25, 4097), (51, 333), (556, 64), (212, 17)])
def test_batchnorm_fwd_bwd(C_l, nvox):
    _lib, ChannelLayout, pad4 = _imports()
    dev = torch.device("cuda")
    lay = ChannelLayout.dense(C_l)
    Cp = lay.C_p
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(C_l, nvox, generator=g, dtype=torch.float64) * 3 + 5).requires_grad_(True)
    gamma = (torch.randn(C_l, generator=g, dtype=torch.float64) + 10).requires_grad_(True)
    beta = torch.randn(C_l, generator=g, dtype=torch.float64).requires_grad_(True)
    rm, rv = torch.zeros(C_l, dtype=torch.float64), torch.ones(C_l, dtype=torch.float64)
    y = F.leaky_relu(F.batch_norm(x.t()[None].transpose(1, 2), rm, rv, gamma, beta, True, 0.1, 1e-5), 0.2)[0]  # (C, nvox)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)

    xcl = to_cl(x.detach().float().to(dev).reshape(C_l, nvox, 1, 1), Cp)
    ws = torch.zeros(int(_lib.lib.dpi_stats_workspace_bytes(Cp)), dtype=torch.uint8, device=dev)
    mp = torch.from_numpy(lay.phys2log()).to(dev)
    gm, bt = gamma.detach().float().to(dev), beta.detach().float().to(dev)
    rmd, rvd = torch.zeros(C_l, device=dev), torch.ones(C_l, device=dev)
    nbt = torch.zeros(1, dtype=torch.int64, device=dev)
    aux = torch.zeros(6, Cp, device=dev)
    ycl = torch.full_like(xcl, 5.0)
    ows = torch.zeros_like(ws)
    _lib.call("dpi_channel_stats", vp(xcl), Cp, nvox, Cp, vp(ws), stream())
    _lib.call("dpi_bn_finalize", vp(ws), nvox, Cp, vp(mp), vp(gm), vp(bt), vp(rmd), vp(rvd), vp(nbt), 0.1, 1e-5,
              vp(aux[0]), vp(aux[1]), vp(aux[2]), vp(aux[3]), stream())
    _lib.call("dpi_affine_act", vp(xcl), Cp, vp(aux[0]), vp(aux[2]), vp(aux[3]), 1, vp(ycl), Cp, nvox, Cp, vp(ows), stream())
    got = ycl[:, :C_l].t().double().cpu()
    assert (got - y.detach()).abs().max().item() <= 2e-5 * y.detach().abs().max().item()
    assert (rmd.double().cpu() - rm).abs().max().item() <= 1e-5 * (1 + rm.abs().max().item())
    assert (rvd.double().cpu() - rv).abs().max().item() <= 1e-5 * (1 + rv.abs().max().item())
    assert int(nbt) == 1
    if Cp > C_l:
        assert ycl[:, C_l:].abs().max().item() == 0.0
    # statistics emitted for the output by the apply pass
    chk = torch.zeros(6, Cp, device=dev)
    _lib.call("dpi_bn_finalize", vp(ows), nvox, Cp, vp(mp), None, None, None, None, None, 0.1, 1e-5, vp(chk[0]), vp(chk[1]),
              vp(chk[2]), vp(chk[3]), stream())
    assert (chk[0, :C_l].double().cpu() - y.detach().mean(1)).abs().max().item() <= 1e-5 * y.detach().abs().max().item()

    dycl = to_cl(dy.float().to(dev).reshape(C_l, nvox, 1, 1), Cp)
    dxcl = torch.full_like(xcl, 1.0)
    dg, db = torch.zeros(C_l, device=dev), torch.zeros(C_l, device=dev)
    # two forms: activation output read from memory, and re-derived from x (out = NULL, scale/shift given)
    for rederive in (False, True):
        dxcl.fill_(1.0)
        dg.zero_(); db.zero_()
        o = None if rederive else ycl
        sc, sh = (aux[2], aux[3]) if rederive else (None, None)
        _lib.call("dpi_bn_bwd_reduce", vp(dycl), Cp, vp(o), Cp, 1, vp(xcl), Cp, vp(aux[0]), vp(aux[1]), vp(sc), vp(sh),
                  nvox, Cp, vp(ws), stream())
        _lib.call("dpi_bn_bwd_finalize", vp(ws), nvox, Cp, vp(mp), vp(dg), vp(db), vp(aux[4]), vp(aux[5]), stream())
        _lib.call("dpi_bn_bwd_apply", vp(dycl), Cp, vp(o), Cp, 1, vp(xcl), Cp, vp(aux[0]), vp(aux[1]), vp(aux[2]), vp(sh),
                  vp(aux[4]), vp(aux[5]), vp(dxcl), Cp, nvox, Cp, 0, stream())
        gdx = dxcl[:, :C_l].t().double().cpu()
        assert (gdx - x.grad).abs().max().item() <= 5e-5 * x.grad.abs().max().item()
        assert (dg.double().cpu() - gamma.grad).abs().max().item() <= 5e-5 * gamma.grad.abs().max().item()
        assert (db.double().cpu() - beta.grad).abs().max().item() <= 5e-5 * beta.grad.abs().max().item()


def test_multi_part_ops_equal_concat_buffer():
    """dpi_*_parts (branch outputs in separate dense buffers) == the same ops on one concatenated buffer, bit for bit"""
    _lib, ChannelLayout, pad4 = _imports()
    dev = torch.device("cuda")
    nvox, widths = 3001, [4, 8, 16]
    Cc = sum(widths)
    g = torch.Generator(device="cpu").manual_seed(7)
    rnd = lambda *sh: torch.randn(*sh, generator=g).to(dev)
    qcat, p, dy = rnd(nvox, Cc), rnd(nvox, Cc), rnd(nvox, Cc)
    qs = [qcat[:, o:o + w].contiguous() for o, w in zip([0, 4, 12], widths)]
    mean, scale, shift, invstd, c1, c2 = (rnd(Cc) for _ in range(6))
    parts = _lib.Parts.make([t.data_ptr() for t in qs], widths, widths)
    ws1 = torch.zeros(int(_lib.lib.dpi_stats_workspace_bytes(Cc)), dtype=torch.uint8, device=dev)
    ws2 = torch.zeros_like(ws1)
    # statistics
    _lib.call("dpi_channel_stats", vp(qcat), Cc, nvox, Cc, vp(ws1), stream())
    _lib.call("dpi_channel_stats_parts", parts, nvox, Cc, vp(ws2), stream())
    assert torch.equal(ws1, ws2)
    # forward add
    y1, y2 = torch.zeros(nvox, Cc, device=dev), torch.zeros(nvox, Cc, device=dev)
    _lib.call("dpi_add_affine_act", vp(p), Cc, vp(qcat), Cc, vp(mean), vp(scale), vp(shift), 1, vp(y1), Cc, nvox, Cc, vp(ws1), stream())
    _lib.call("dpi_add_affine_act_parts", vp(p), Cc, parts, vp(mean), vp(scale), vp(shift), 1, vp(y2), Cc, nvox, Cc, vp(ws2), stream())
    assert torch.equal(y1, y2) and torch.equal(ws1, ws2)
    # backward reduce / apply (part 1 accumulates)
    _lib.call("dpi_bn_bwd_reduce", vp(dy), Cc, vp(y1), Cc, 1, vp(qcat), Cc, vp(mean), vp(invstd), None, None, nvox, Cc, vp(ws1), stream())
    _lib.call("dpi_bn_bwd_reduce_parts", vp(dy), Cc, vp(y1), Cc, 1, parts, vp(mean), vp(invstd), nvox, Cc, vp(ws2), stream())
    assert torch.equal(ws1, ws2)
    dx_cat = torch.full((nvox, Cc), 0.5, device=dev)
    dxs = [torch.full((nvox, w), 0.5, device=dev) for w in widths]
    dparts = _lib.Parts.make([t.data_ptr() for t in dxs], widths, widths)
    for acc in (0, 1):
        _lib.call("dpi_bn_bwd_apply", vp(dy), Cc, vp(y1), Cc, 1, vp(qcat), Cc, vp(mean), vp(invstd), vp(scale), None, vp(c1), vp(c2),
                  vp(dx_cat), Cc, nvox, Cc, acc, stream())
        _lib.call("dpi_bn_bwd_apply_parts", vp(dy), Cc, vp(y1), Cc, 1, parts, vp(mean), vp(invstd), vp(scale), vp(c1), vp(c2),
                  dparts, 7 if acc else 0, None, 0, nvox, Cc, stream())
        assert torch.equal(dx_cat, torch.cat(dxs, 1)), "accumulate=%d" % acc
    # mixed accumulate mask: only part 1 accumulates
    before = [t.clone() for t in dxs]
    dp_fused = torch.full((nvox, Cc), 9.0, device=dev)
    _lib.call("dpi_bn_bwd_apply_parts", vp(dy), Cc, vp(y1), Cc, 1, parts, vp(mean), vp(invstd), vp(scale), vp(c1), vp(c2),
              dparts, 2, vp(dp_fused), Cc, nvox, Cc, stream())
    dp_ref = torch.zeros(nvox, Cc, device=dev)
    _lib.call("dpi_act_bwd", vp(dy), Cc, vp(y1), Cc, 1, vp(dp_ref), Cc, nvox, Cc, 0, stream())
    assert torch.equal(dp_fused, dp_ref), "fused second output dp = dy * act'(out)"
    fresh = torch.zeros(nvox, Cc, device=dev)
    _lib.call("dpi_bn_bwd_apply", vp(dy), Cc, vp(y1), Cc, 1, vp(qcat), Cc, vp(mean), vp(invstd), vp(scale), None, vp(c1), vp(c2),
              vp(fresh), Cc, nvox, Cc, 0, stream())
    assert torch.equal(dxs[0], fresh[:, 0:4]) and torch.equal(dxs[2], fresh[:, 12:28])
    assert torch.allclose(dxs[1], before[1] + fresh[:, 4:12], rtol=1e-6, atol=1e-6)


def test_bn_finalize_over_per_part_statistics():
    """dpi_bn_finalize_parts over three workspaces (the statistics each branch's own pass left) == dpi_bn_finalize over the
    statistics of the concatenation; different producers -> different numbers of partial rows per part"""
    _lib, ChannelLayout, pad4 = _imports()
    dev = torch.device("cuda")
    nvox, widths = 70001, [4, 8, 16]
    Cc = sum(widths)
    g = torch.Generator(device="cpu").manual_seed(5)
    qs = [(torch.randn(nvox, w, generator=g) * 2 + 1).to(dev) for w in widths]
    parts = _lib.Parts.make([q.data_ptr() for q in qs], widths, widths)
    wsz = lambda c: int(_lib.lib.dpi_stats_workspace_bytes(c))
    ws_cat = torch.zeros(wsz(Cc), dtype=torch.uint8, device=dev)
    _lib.call("dpi_channel_stats_parts", parts, nvox, Cc, vp(ws_cat), stream())
    wss = [torch.zeros(wsz(w), dtype=torch.uint8, device=dev) for w in widths]
    ys = [torch.zeros_like(q) for q in qs]
    for q, y, w, ws in zip(qs, ys, widths, wss):
        # identity affine pass that leaves the statistics of its output (what BnActOp does with emit_stats)
        _lib.call("dpi_affine_act", vp(q), w, None, None, None, 0, vp(y), w, nvox, w, vp(ws), stream())
        assert torch.equal(q, y)
    gm, bt = torch.randn(Cc, generator=g).to(dev), torch.randn(Cc, generator=g).to(dev)
    outs = []
    for which in (0, 1):
        rm, rv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
        nbt = torch.zeros(1, dtype=torch.int64, device=dev)
        aux = torch.zeros(4, Cc, device=dev)
        if which == 0:
            _lib.call("dpi_bn_finalize", vp(ws_cat), nvox, Cc, None, vp(gm), vp(bt), vp(rm), vp(rv), vp(nbt), 0.1, 1e-5,
                      vp(aux[0]), vp(aux[1]), vp(aux[2]), vp(aux[3]), stream())
        else:
            sp = _lib.StatsParts.make([w.data_ptr() for w in wss], widths)
            _lib.call("dpi_bn_finalize_parts", sp, nvox, Cc, None, vp(gm), vp(bt), vp(rm), vp(rv), vp(nbt), 0.1, 1e-5,
                      vp(aux[0]), vp(aux[1]), vp(aux[2]), vp(aux[3]), stream())
        assert int(nbt) == 1
        outs.append((aux.clone(), rm.clone(), rv.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    ref = torch.cat(qs, 1).double()
    assert torch.allclose(outs[1][0][0].double(), ref.mean(0), rtol=1e-6, atol=1e-6)


def _ws_rows(ws, C):
    hdr = ws[:16].view(torch.int64).cpu()
    n = int(hdr[0])
    assert int(hdr[1]) == C and 1 <= n <= 592
    return ws[16:16 + n * 2 * C * 8].view(torch.float64).view(n, 2, C).sum(0).cpu()


@pytest.mark.parametrize("nvox", [3001, 70000])
def test_bn_bwd_apply_with_the_next_reduce_fused(nvox):
    """dpi_bn_bwd_apply_next / dpi_bn_bwd_apply_parts_next == the plain apply pass (bit for bit) followed by the separate
    reduce pass of the next unit (sums equal up to the order of the fp32 pre-sums), for both kinds of dpi_bn_next_reduce"""
    _lib, ChannelLayout, pad4 = _imports()
    dev = torch.device("cuda")
    widths = [4, 8, 16]
    Cc = sum(widths)
    g = torch.Generator(device="cpu").manual_seed(11)
    rnd = lambda *sh: torch.randn(*sh, generator=g).to(dev)
    wsz = int(_lib.lib.dpi_stats_workspace_bytes(Cc))
    ws_ref, ws_f = torch.zeros(wsz, dtype=torch.uint8, device=dev), torch.zeros(wsz, dtype=torch.uint8, device=dev)
    mean, invstd, scale, shift, c1, c2 = (rnd(Cc) for _ in range(6))
    nmean, ninv, nscale, nshift = (rnd(Cc) for _ in range(4))
    dy = rnd(nvox, Cc)

    # kind 1: norm2 (no activation of its own) on t = act(...); the next unit's x is the concatenated branches
    t = torch.nn.functional.leaky_relu(rnd(nvox, Cc), 0.2)
    qs = [rnd(nvox, w) for w in widths]
    parts = _lib.Parts.make([q.data_ptr() for q in qs], widths, widths)
    for sh in (None, shift):                   # this unit without / with an activation re-derived from x
        act = 0 if sh is None else 1
        dx_ref, dx_f = torch.full((nvox, Cc), 3.0, device=dev), torch.full((nvox, Cc), 4.0, device=dev)
        _lib.call("dpi_bn_bwd_apply", vp(dy), Cc, None, Cc, act, vp(t), Cc, vp(mean), vp(invstd), vp(scale), vp(sh), vp(c1),
                  vp(c2), vp(dx_ref), Cc, nvox, Cc, 0, stream())
        _lib.call("dpi_bn_bwd_reduce_parts", vp(dx_ref), Cc, vp(t), Cc, 1, parts, vp(nmean), vp(ninv), nvox, Cc, vp(ws_ref),
                  stream())
        nx = _lib.NextReduce.make(1, 1, parts, nmean.data_ptr(), ninv.data_ptr(), 0, 0, ws_f.data_ptr())
        _lib.call("dpi_bn_bwd_apply_next", vp(dy), Cc, None, Cc, act, vp(t), Cc, vp(mean), vp(invstd), vp(scale), vp(sh),
                  vp(c1), vp(c2), vp(dx_f), Cc, nvox, Cc, 0, nx, stream())
        assert torch.equal(dx_ref, dx_f)
        r_ref, r_f = _ws_rows(ws_ref, Cc), _ws_rows(ws_f, Cc)
        assert (r_ref - r_f).abs().max().item() <= 2e-6 * r_ref.abs().max().item()

    # kind 2: the apply pass of act(p + BN(cat)) with the second output dp; the next unit is conv + BN [+ act] with x'
    y = torch.nn.functional.leaky_relu(rnd(nvox, Cc), 0.2)
    xs = rnd(nvox, Cc)
    xparts = _lib.Parts.make([xs.data_ptr()], [Cc], [Cc])
    for mask in (0, 2):
        for nact in (0, 1):
            dxr = [torch.full((nvox, w), 0.5, device=dev) for w in widths]
            dxf = [torch.full((nvox, w), 0.5, device=dev) for w in widths]
            dpr, dpf = torch.zeros(nvox, Cc, device=dev), torch.ones(nvox, Cc, device=dev)
            pr = _lib.Parts.make([d.data_ptr() for d in dxr], widths, widths)
            pf = _lib.Parts.make([d.data_ptr() for d in dxf], widths, widths)
            _lib.call("dpi_bn_bwd_apply_parts", vp(dy), Cc, vp(y), Cc, 1, parts, vp(mean), vp(invstd), vp(scale), vp(c1), vp(c2),
                      pr, mask, vp(dpr), Cc, nvox, Cc, stream())
            _lib.call("dpi_bn_bwd_reduce", vp(dpr), Cc, None, Cc, nact, vp(xs), Cc, vp(nmean), vp(ninv),
                      vp(nscale if nact else None), vp(nshift if nact else None), nvox, Cc, vp(ws_ref), stream())
            nx = _lib.NextReduce.make(2, nact, xparts, nmean.data_ptr(), ninv.data_ptr(), nscale.data_ptr() if nact else 0,
                                      nshift.data_ptr() if nact else 0, ws_f.data_ptr())
            _lib.call("dpi_bn_bwd_apply_parts_next", vp(dy), Cc, vp(y), Cc, 1, parts, vp(mean), vp(invstd), vp(scale), vp(c1),
                      vp(c2), pf, mask, vp(dpf), Cc, nvox, Cc, nx, stream())
            assert torch.equal(dpr, dpf) and all(torch.equal(a, b) for a, b in zip(dxr, dxf))
            r_ref, r_f = _ws_rows(ws_ref, Cc), _ws_rows(ws_f, Cc)
            assert (r_ref - r_f).abs().max().item() <= 2e-6 * r_ref.abs().max().item(), (mask, nact)
    # kind 3: the next unit is an add without BatchNorm - both addends receive dx * act'(x), nothing else is written
    for sh in (None, shift):
        act = 0 if sh is None else 1
        dx_ref, g_ref = torch.zeros(nvox, Cc, device=dev), torch.zeros(nvox, Cc, device=dev)
        _lib.call("dpi_bn_bwd_apply", vp(dy), Cc, None, Cc, act | 0x100, vp(t), Cc, vp(mean), vp(invstd), vp(scale), vp(sh),
                  vp(c1), vp(c2), vp(dx_ref), Cc, nvox, Cc, 0, stream())
        _lib.call("dpi_act_bwd", vp(dx_ref), Cc, vp(t), Cc, 1, vp(g_ref), Cc, nvox, Cc, 0, stream())
        gq, gp = torch.full((nvox, Cc), 7.0, device=dev), torch.full((nvox, Cc), 8.0, device=dev)
        nx = _lib.NextReduce.make(3, 1, _lib.Parts.make([gp.data_ptr()], [Cc], [Cc]), 0, 0, 0, 0, 0)
        _lib.call("dpi_bn_bwd_apply_next", vp(dy), Cc, None, Cc, act | 0x100, vp(t), Cc, vp(mean), vp(invstd), vp(scale),
                  vp(sh), vp(c1), vp(c2), vp(gq), Cc, nvox, Cc, 0, nx, stream())
        assert torch.equal(gq, g_ref) and torch.equal(gp, g_ref)
    # a combination that cannot carry the fused reduce is refused, not silently mis-computed
    nx = _lib.NextReduce.make(1, 1, parts, nmean.data_ptr(), ninv.data_ptr(), 0, 0, ws_f.data_ptr())
    with pytest.raises(_lib.DpiError):
        _lib.call("dpi_bn_bwd_apply_next", vp(dy), Cc, None, Cc, 0, vp(t), Cc, vp(mean), vp(invstd), vp(scale), None, vp(c1),
                  vp(c2), vp(dx_f), Cc, nvox, Cc, 1, nx, stream())


@pytest.mark.parametrize("mode", ["nearest", "linear"])
@pytest.mark.parametrize("dims,odims,up_d", [((4, 5, 6), (8, 10, 12), 1), ((3, 4, 5), (5, 7, 9), 1), ((1, 11, 7), (1, 22, 13), 0),
                                             # >= 64 K input voxels: the several-rows-per-CTA backward kernel, odd sizes
                                             # (partial row blocks) with and without the centre crop
                                             ((33, 47, 45), (65, 93, 89), 1), ((34, 45, 44), (68, 90, 88), 1)])
def test_upsample_fwd_bwd(mode, dims, odims, up_d):
    _lib, ChannelLayout, pad4 = _imports()
    dev = torch.device("cuda")
    Cc = 8
    g = torch.Generator().manual_seed(3)
    x = torch.randn((Cc,) + dims, generator=g, dtype=torch.float64, requires_grad=True)
    if up_d:
        full = F.interpolate(x[None], scale_factor=2, mode="nearest" if mode == "nearest" else "trilinear")[0]
    else:
        full = F.interpolate(x[None, :, 0], scale_factor=2, mode="nearest" if mode == "nearest" else "bilinear")[0][:, None]
    y = full[:, :odims[0], :odims[1], :odims[2]]
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dy)
    xcl = to_cl(x.detach().float().to(dev), Cc)
    ycl = torch.zeros((int(np.prod(odims)), 16), device=dev)   # write into a slice at channel offset 4
    m = 0 if mode == "nearest" else 1
    yptr = C.c_void_p(ycl.data_ptr() + 16)
    _lib.call("dpi_upsample2x_fwd", vp(xcl), Cc, *dims, yptr, 16, *odims, Cc, m, up_d, stream())
    got = from_cl(ycl[:, 4:12], Cc, odims).double().cpu()
    assert (got - y.detach()).abs().max().item() <= 1e-6
    assert ycl[:, :4].abs().max().item() == 0 and ycl[:, 12:].abs().max().item() == 0
    dycl = torch.zeros_like(ycl)
    dycl[:, 4:12] = to_cl(dy.float().to(dev), Cc)
    dxcl = torch.zeros_like(xcl)
    _lib.call("dpi_upsample2x_bwd", C.c_void_p(dycl.data_ptr() + 16), 16, *odims, vp(dxcl), Cc, *dims, Cc, m, up_d, 0, stream())
    gdx = from_cl(dxcl, Cc, dims).double().cpu()
    assert (gdx - x.grad).abs().max().item() <= 1e-5
    # accumulate: bit-identical to "old value + the fresh result"
    fresh = dxcl.clone()
    dxcl.fill_(0.25)
    _lib.call("dpi_upsample2x_bwd", C.c_void_p(dycl.data_ptr() + 16), 16, *odims, vp(dxcl), Cc, *dims, Cc, m, up_d, 1, stream())
    assert torch.equal(dxcl, fresh + 0.25)


@pytest.mark.parametrize("kind", ["mae", "mse"])
def test_masked_loss_and_metrics(kind):
    _lib, ChannelLayout, pad4 = _imports()
    dev = torch.device("cuda")
    n = 4 * 12345
    g = torch.Generator().manual_seed(5)
    out = torch.randn(n, generator=g, dtype=torch.float64, requires_grad=True)
    img = torch.randn(n, generator=g, dtype=torch.float64) * 2
    mask = (torch.rand(n, generator=g) > 0.6).double()
    a, b = out * mask, img * mask
    loss = F.mse_loss(a, b) if kind == "mse" else F.l1_loss(a, b)
    loss.backward()
    snr = 10 * torch.log10((img ** 2).sum() / ((img - out.detach()) ** 2).sum())
    td, od = img - img.mean(), out.detach() - out.detach().mean()
    pc = (td * od).sum() / (td.pow(2).sum().sqrt() * od.pow(2).sum().sqrt())
    o, i_, m_ = out.detach().float().to(dev), img.float().to(dev), mask.float().to(dev)
    dout = torch.zeros(n, device=dev)
    ws = torch.zeros(int(_lib.lib.dpi_loss_workspace_bytes()), dtype=torch.uint8, device=dev)
    sc = torch.zeros(8, dtype=torch.float64, device=dev)
    _lib.call("dpi_masked_loss", vp(o), vp(i_), vp(m_), n, n, _lib.LOSS_CODES[kind], vp(dout), vp(ws), ws.numel(), vp(sc), stream())
    sc = sc.cpu()
    assert abs(sc[0].item() - loss.item()) <= 1e-6 * abs(loss.item())
    assert abs(sc[1].item() - snr.item()) <= 1e-4
    assert abs(sc[2].item() - pc.item()) <= 1e-5
    assert (dout.double().cpu() - out.grad).abs().max().item() <= 1e-6 * out.grad.abs().max().item()


def test_adam_matches_torch():
    _lib, ChannelLayout, pad4 = _imports()
    dev = torch.device("cuda")
    n = 10007
    torch.manual_seed(0)
    p0 = torch.randn(n)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-3)
    p, m, v = p0.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    for step in range(1, 6):
        g = torch.randn(n) * (0.1 ** step)
        p_ref.grad = g.clone()
        opt.step()
        gd = g.to(dev)
        _lib.call("dpi_adam_step", vp(p), vp(gd), vp(m), vp(v), n, 1e-3, 0.9, 0.999, 1e-8, 0.0, step, stream())
        diff = (p.cpu() - p_ref.detach()).abs().max().item()
        assert diff <= 2e-7, (step, diff)


def test_noise_statistics_and_axpy():
    _lib, ChannelLayout, pad4 = _imports()
    dev = torch.device("cuda")
    n = 1 << 22
    z = torch.zeros(n, device=dev)
    out = torch.empty(n, device=dev)
    _lib.call("dpi_noise_axpy", vp(z), None, vp(out), n, 1.0, 42, 0, 0, stream())
    assert abs(out.mean().item()) < 3e-3 and abs(out.std().item() - 1) < 3e-3
    assert abs((out ** 3).mean().item()) < 2e-2 and abs((out ** 4).mean().item() - 3) < 5e-2
    out2 = torch.empty(n, device=dev)
    _lib.call("dpi_noise_axpy", vp(z), None, vp(out2), n, 1.0, 42, 1, 0, stream())
    assert abs((out * out2).mean().item()) < 3e-3      # different offsets decorrelate
    out3 = torch.empty(n, device=dev)
    _lib.call("dpi_noise_axpy", vp(z), None, vp(out3), n, 1.0, 42, 0, 0, stream())
    assert torch.equal(out, out3)                        # same (seed, offset) -> same stream
    eps = torch.randn(n, device=dev)
    zz = torch.randn(n, device=dev)
    _lib.call("dpi_noise_axpy", vp(zz), vp(eps), vp(out), n, 0.03, 0, 0, 0, stream())
    assert (out - (zz + 0.03 * eps)).abs().max().item() <= 1e-6


def test_layout_roundtrip():
    _lib, ChannelLayout, pad4 = _imports()
    dev = torch.device("cuda")
    lay = ChannelLayout.concat([ChannelLayout.dense(4), ChannelLayout.dense(8), ChannelLayout.dense(13)])
    nvox = 1001
    x = torch.randn(lay.C_l, nvox, device=dev)
    mp = torch.from_numpy(lay.phys2log()).to(dev)
    cl = torch.full((nvox, lay.C_p), 5.0, device=dev)
    _lib.call("dpi_nchw_to_cl", vp(x), lay.C_l, nvox, vp(mp), vp(cl), lay.C_p, lay.C_p, stream())
    for p in range(lay.C_p):
        l = int(mp[p])
        assert torch.equal(cl[:, p], x[l] if l >= 0 else torch.zeros(nvox, device=dev))
    back = torch.zeros_like(x)
    _lib.call("dpi_cl_to_nchw", vp(cl), lay.C_p, lay.C_p, vp(mp), vp(back), lay.C_l, nvox, stream())
    assert torch.equal(back, x)


FUSED_DGRAD_CASES = [
    # (dims, Cin, Cout_a (3x3(x3), all taps), Cout_b (1x1), kd)
    ((12, 16, 8), 67, 4, 25, 3),        # decoder block of level 0 (3.conv3x3 + 3.shortcut): plain march, N = 72
    ((9, 33, 17), 25, 16, 16, 3),       # ResPath of level 0 (conv3x3 + conv1x1): packed march, N = 28
    ((10, 20, 12), 137, 8, 51, 3),      # decoder block of level 1: N = 140 > 128, falls back to kernels that ignore the hint
    ((6, 16, 16), 25, 8, 51, 3),        # encoder block of level 1: two channel chunks (8 + 52 = 60)
    ((1, 40, 30), 25, 8, 26, 1),        # 2-D block
]


@pytest.mark.parametrize("ctas", ["3", "148"])
@pytest.mark.parametrize("case", FUSED_DGRAD_CASES)
def test_conv_dgrad_fused(case, ctas, monkeypatch):
    """dpi_conv_dgrad_fused == dgrad(3x3 conv) + dgrad(1x1 conv) of two convs that share their input
    (mulresunet.py:85-92,108-109), and the batched weight pack writes the combined transposed weights"""
    monkeypatch.setenv("DPI_TC_MARCH_CTAS", ctas)
    _lib, ChannelLayout, pad4 = _imports()
    dims, cin, ca, cb, kd = case
    k = (kd, 3, 3)
    taps = kd * 9
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(77)
    x = torch.randn((cin,) + dims, generator=g, dtype=torch.float64, requires_grad=True)
    wa = torch.randn((ca, cin) + k, generator=g, dtype=torch.float64) * 0.1
    wb = torch.randn((cb, cin, 1, 1, 1), generator=g, dtype=torch.float64) * 0.1
    ya = F.conv3d(x[None], wa, None, padding=(kd // 2, 1, 1))[0]
    yb = F.conv3d(x[None], wb, None)[0]
    dya = torch.randn(ya.shape, generator=g, dtype=torch.float64)
    dyb = torch.randn(yb.shape, generator=g, dtype=torch.float64)
    (ya * dya).sum().backward(retain_graph=True)
    dxa = x.grad.clone()
    x.grad = None
    (yb * dyb).sum().backward()
    dx_ref = dxa + x.grad
    cip, cap, cbp = pad4(cin), pad4(ca), pad4(cb)
    cc = cap + cbp
    nvox = int(np.prod(dims))
    dycat = torch.zeros((nvox, cc), device=dev)
    dycat[:, :ca] = dya.float().to(dev).reshape(ca, -1).t()
    dycat[:, cap:cap + cb] = dyb.float().to(dev).reshape(cb, -1).t()
    _, wda = pack_w(wa.float().to(dev), cap, cip)             # [cip][taps][cap]
    _, wdb = pack_w(wb.float().to(dev), cbp, cip)             # [cip][1][cbp]
    wt = torch.zeros((cip, taps, cc), device=dev)
    wt[:, :, :cap] = wda
    wt[:, taps // 2, cap:] = wdb[:, 0, :]
    wt = wt.contiguous()
    geom = _lib.ConvGeom(dims[0], dims[1], dims[2], cip, cc, kd, 3, 3, 1)
    for prec, tol in ((1, 3e-3), (0, 2e-5)):
        dx = torch.full((nvox, cip), 3.0, device=dev)
        _lib.call("dpi_conv_dgrad_fused", vp(dycat), cc, vp(wt), vp(dx), cip, C.byref(geom), cap, 0, prec, stream())
        got = from_cl(dx, cin, dims).double().cpu()
        sc = dx_ref.abs().max().item()
        assert (got - dx_ref).abs().max().item() <= tol * sc, ("fused dgrad", prec)
        if cip > cin:
            assert dx[:, cin:].abs().max().item() == 0.0
        _lib.call("dpi_conv_dgrad_fused", vp(dycat), cc, vp(wt), vp(dx), cip, C.byref(geom), cap, 1, prec, stream())
        got2 = from_cl(dx, cin, dims).double().cpu()
        assert (got2 - 2 * dx_ref).abs().max().item() <= 2 * tol * sc, ("fused dgrad accumulate", prec)
    # the hint only skips work: the plain data gradient of the same 'wide' conv gives the same values bit for bit
    dx1 = torch.zeros((nvox, cip), device=dev)
    dx2 = torch.zeros((nvox, cip), device=dev)
    _lib.call("dpi_conv_dgrad_fused", vp(dycat), cc, vp(wt), vp(dx1), cip, C.byref(geom), cap, 0, 1, stream())
    _lib.call("dpi_conv_dgrad", vp(dycat), cc, vp(wt), vp(dx2), cip, C.byref(geom), 0, 1, stream())
    assert (dx1 - dx2).abs().max().item() <= 1e-5 * dx2.abs().max().item()

    # batched pack: both layers write their slice of the combined transposed weights
    mo_a = torch.arange(cap, dtype=torch.int32, device=dev)
    mo_a[ca:] = -1
    mo_b = torch.arange(cbp, dtype=torch.int32, device=dev)
    mo_b[cb:] = -1
    mi = torch.arange(cip, dtype=torch.int32, device=dev)
    mi[cin:] = -1
    wt2 = torch.zeros_like(wt)
    wa32, wb32 = wa.float().to(dev).contiguous(), wb.float().to(dev).contiguous()
    jobs = (_lib.PackJob * 2)(
        _lib.PackJob(wa32.data_ptr(), None, None, None, None, None, None, mo_a.data_ptr(), mi.data_ptr(), ca, cin, cap, cip,
                     taps, 0, wt2.data_ptr(), cc, 0, taps, 0),
        _lib.PackJob(wb32.data_ptr(), None, None, None, None, None, None, mo_b.data_ptr(), mi.data_ptr(), cb, cin, cbp, cip,
                     1, 0, wt2.data_ptr(), cc, cap, taps, taps // 2))
    jd = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(dev)
    _lib.call("dpi_pack_conv_weights_batched", vp(jd), 2, 0, stream())
    assert torch.equal(wt2, wt)
