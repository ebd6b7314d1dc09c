import sys, ctypes as C, torch
sys.path.insert(0, '.')
from deep_prior_interpolation_b200 import _lib
dev = torch.device("cuda"); torch.cuda.set_device(0)
vp = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
nvox = 256 * 128 * 128
for Cc in (4, 8, 16, 28):
    nset = max(3, int(600e6 // (nvox * Cc * 4 * 2)) + 1)
    sets = [(torch.randn(nvox, Cc, device=dev), torch.randn(nvox, Cc, device=dev), torch.empty(nvox, Cc, device=dev)) for _ in range(nset)]
    aux = torch.rand(6, Cc, device=dev) + 0.5
    ws = torch.zeros(int(_lib.lib.dpi_stats_workspace_bytes(Cc)), dtype=torch.uint8, device=dev)
    def reduce(i):
        dy, x, dx = sets[i % nset]
        _lib.call("dpi_bn_bwd_reduce", vp(dy), Cc, None, Cc, 1, vp(x), Cc, vp(aux[0]), vp(aux[1]), vp(aux[2]), vp(aux[3]), nvox, Cc, vp(ws), st)
    def apply(i):
        dy, x, dx = sets[i % nset]
        _lib.call("dpi_bn_bwd_apply", vp(dy), Cc, None, Cc, 1, vp(x), Cc, vp(aux[0]), vp(aux[1]), vp(aux[2]), vp(aux[3]), vp(aux[4]), vp(aux[5]), vp(dx), Cc, nvox, Cc, 0, st)
    def affine(i):
        dy, x, dx = sets[i % nset]
        _lib.call("dpi_affine_act", vp(x), Cc, vp(aux[0]), vp(aux[2]), vp(aux[3]), 1, vp(dx), Cc, nvox, Cc, None, st)
    def stats(i):
        dy, x, dx = sets[i % nset]
        _lib.call("dpi_channel_stats", vp(x), Cc, nvox, Cc, vp(ws), st)
    def copy(i):
        dy, x, dx = sets[i % nset]
        dx.copy_(x)
    for name, fn, passes in (("reduce", reduce, 2), ("apply", apply, 3), ("affine_act", affine, 2), ("stats", stats, 1), ("torch copy", copy, 2)):
        for i in range(3): fn(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 30
        e0.record()
        for i in range(n): fn(i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        print("C=%2d %-10s %7.1f us  %5.2f TB/s" % (Cc, name, us, passes * nvox * Cc * 4 / us / 1e6))
    del sets
