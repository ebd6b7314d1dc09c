"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run these sizes in
seconds): the config-4 patch grid (1000x256x256 volume, 64^3 patches at 50 % stride = 1470 patches) and one
(256,128,128) patch of the default MulResUnet3D through the whole iteration (bench.py's workload)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_config4_patch_grid_full_size():
    """extract -> reassemble at the size of BASELINE.json configs[3] (data.py:44-130, patch_extractor.py:299-428).

    Integer-valued data make every property exact: each patch equals its NumPy window (spot-checked), the sum over all
    patches equals sum(volume x hit count) (a checksum over all 385 M gathered values), and the overlap-average of
    untouched patches returns the cropped volume bit for bit, also through the gain division."""
    from deep_prior_interpolation_b200.data import PatchExtractor
    rng = np.random.RandomState(4)
    shape, dim, stride = (1000, 256, 256), (64, 64, 64), (32, 32, 32)
    vol = rng.randint(-1000, 1000, size=shape).astype(np.float64)
    pe = PatchExtractor(dim=dim, stride=stride)
    pa = pe.extract(vol)
    assert pa.shape == (30, 7, 7, 64, 64, 64) and pa.dtype == np.float64          # 1470 patches
    assert tuple(pe.in_content_cropped_shape) == (992, 256, 256)                  # last 8 time samples dropped
    for _ in range(24):
        i, j, k = rng.randint(30), rng.randint(7), rng.randint(7)
        win = vol[32 * i:32 * i + 64, 32 * j:32 * j + 64, 32 * k:32 * k + 64]
        assert np.array_equal(pa[i, j, k], win), (i, j, k)
    cnt = [np.zeros(n) for n in (992, 256, 256)]
    for a, (n, npatch) in enumerate(((992, 30), (256, 7), (256, 7))):
        for p in range(npatch):
            cnt[a][32 * p:32 * p + 64] += 1
    hits = cnt[0][:, None, None] * cnt[1][None, :, None] * cnt[2][None, None, :]
    assert float(pa.sum()) == float((vol[:992] * hits).sum())                     # integers < 2^53: exact
    f = pa.astype(np.float32)
    del pa
    rec = pe.reconstruct(f)
    assert rec.dtype == np.float32 and rec.shape == (992, 256, 256)
    assert np.array_equal(rec, vol[:992].astype(np.float32))
    rec40 = pe.reconstruct(f, gain=40.0)
    assert np.array_equal(rec40, vol[:992].astype(np.float32) / np.float32(40.0))


def test_full_size_patch_iteration_properties():
    """One (256,128,128) patch, default widths, TF32 tcgen05 path, CUDA-graph replay (bench.py's step):
    * bit-reproducible: the same 24 iterations from the same weights, noise seed and Adam state give the same
      {loss, snr, pcorr} history and the same parameters, bit for bit (no atomics anywhere on the path);
    * the masked loss / SNR / PCORR the device reports equal main.py:161-167 evaluated with torch (fp64) on the
      network output the device holds (4.2 M voxels);
    * BatchNorm bookkeeping: every num_batches_tracked advanced by the number of iterations;
    * the loss comes down."""
    import bench
    from deep_prior_interpolation_b200.interpolator import Interpolator
    dev = torch.device("cuda", 0)
    dims = (256, 128, 128)
    args = bench.default_args("tf32")
    args.epochs = 32
    img_np, mask_np = bench.synthetic_patch(dims, seed=1)
    T = Interpolator(args, outpath="/tmp")
    T.load_data({"image": img_np, "mask": mask_np, "name": "0"})
    T.build_model()
    T.build_input()
    eng = T.net.engine_for(dims, dev, max_iters=args.epochs)
    eng.set_loss("mae")
    eng.set_noise_input(T.input_)
    eng.set_target(T.img_, T.mask_)
    eng.reset_loop_state(1e-3, 0)
    eng.capture(0.03, 0)
    P0, B0, I0 = eng.params.P.clone(), eng.params.B.clone(), eng.params.I.clone()
    n = 24

    def run():
        eng.params.P.copy_(P0)
        eng.params.B.copy_(B0)
        eng.params.I.copy_(I0)
        eng.reset_loop_state(1e-3, 0)
        for _ in range(n):
            eng.graph.replay()
        torch.cuda.synchronize()
        return eng.history[:n].clone(), eng.params.P.clone(), eng.params.I.clone()

    h1, p1, i1 = run()
    h2, p2, i2 = run()
    assert torch.isfinite(h1).all()
    assert torch.equal(h1, h2) and torch.equal(p1, p2), "graph replay must be bit-reproducible"
    assert torch.equal(i1 - I0, torch.full_like(I0, n)), "num_batches_tracked of every BatchNorm"
    # (the first iterations are erratic - BatchNorm statistics of a freshly initialised net - so compare the tail)
    assert float(h1[n - 8:, 0].min()) < float(h1[0, 0]), h1[:, 0]

    # the device-side loss / metrics of the last iteration against torch on the same output
    out = eng.output_nchw().double()
    img, mask = T.img_.to(dev).double(), T.mask_.to(dev).double()
    loss = (out * mask - img * mask).abs().mean()
    snr = 10 * torch.log10((img ** 2).sum() / ((img - out) ** 2).sum())
    td, od = img - img.mean(), out - out.mean()
    pc = (td * od).sum() / ((td ** 2).sum().sqrt() * (od ** 2).sum().sqrt())
    l, s, p = (float(v) for v in h1[n - 1, :3])
    assert abs(l - float(loss)) <= 1e-6 * abs(float(loss)), (l, float(loss))
    assert abs(s - float(snr)) <= 1e-5 and abs(p - float(pc)) <= 1e-6, ((s, float(snr)), (p, float(pc)))
