"""Patch extraction / reassembly with the semantics of the reference's ``data.py`` and
``utils/patch_extractor.py``, executed by the bit-exact index kernels ``dpi_patch_extract_f64`` /
``dpi_patch_reassemble_f32`` (``include/dpi_b200.h``).

* ``extract_patches(args)``  — ``data.py:44-84``: load the two ``.npy`` volumes, NaN traces -> binary mask
  (``utils/processing.py:27-31``), strided window gather (``patch_extractor.py:299-362``: ``(n-p)//s+1`` windows
  per axis, tail dropped, C order), 2.5-D transposes (``data.py:20-41``), ``* gain``.
* ``reconstruct_patches(args)`` — ``data.py:87-130``: gather the per-patch ``*_run.npy`` results (sorted — the
  reference relies on file-system order, ``data.py:99``), inverse transposes, float64 overlap-add in patch order,
  divide by the hit count, cast to float32, ``/ gain`` (``patch_extractor.py:370-428``, ``data.py:116``).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import sys
import types
from glob import glob
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import utils as u

__all__ = ["PatchExtractor", "extract_patches", "reconstruct_patches", "load_run", "history_alias", "patch_array_shape", "count_patches",
           "in_content_cropped_shape"]


def patch_array_shape(in_size, patch_size, patch_stride) -> tuple:
    """``patch_array_shape`` (patch_extractor.py:153-155)"""
    idx = (np.array(in_size) - np.array(patch_size)) // np.array(patch_stride) + 1
    return tuple(int(i) for i in idx) + tuple(patch_size)


def count_patches(in_size, patch_size, patch_stride) -> int:
    """``count_patches`` (patch_extractor.py:140-150)"""
    return int(np.prod((np.array(in_size) - np.array(patch_size)) // np.array(patch_stride) + 1))


def in_content_cropped_shape(in_size, patch_size, patch_stride) -> tuple:
    n = len(in_size)
    idx = patch_array_shape(in_size, patch_size, patch_stride)[:n]
    return tuple(int((idx[a] - 1) * patch_stride[a] + patch_size[a]) for a in range(n))


def _i32x3(v: Sequence[int]):
    return (C.c_int32 * 3)(*[int(x) for x in v])


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("patch kernels need a CUDA device; there is no CPU fallback in the product path")
    return torch.device("cuda", torch.cuda.current_device())


class PatchExtractor:
    """N-D (N<=3) sliding-window extractor / overlap-averaging reconstructor (rect taper, zero offset) —
    the subset of ``utils/patch_extractor.py:PatchExtractor`` that ``data.py`` uses."""

    def __init__(self, dim: tuple, stride: tuple = None):
        if not isinstance(dim, tuple):
            raise ValueError("dim must be a tuple")
        self.dim = tuple(int(d) for d in dim)
        self.ndim = len(dim)
        if self.ndim > 3:
            raise ValueError("PatchExtractor handles up to 3 dimensions")
        stride = self.dim if stride is None else stride
        if not isinstance(stride, tuple) or len(stride) != self.ndim:
            raise ValueError("stride must a tuple of length %d" % self.ndim)
        self.stride = tuple(int(s) for s in stride)
        self.in_content_original_shape = None
        self.in_content_cropped_shape = None
        self.patch_array_shape = None

    def _pad3(self, v, fill=1):
        return (fill,) * (3 - self.ndim) + tuple(v)

    def extract(self, in_content: np.ndarray, gain: float = 1.0) -> np.ndarray:
        if not isinstance(in_content, np.ndarray):
            raise ValueError("in_content must be of type: " + str(np.ndarray))
        if in_content.ndim != self.ndim:
            raise ValueError("in_content shape must a tuple of length %d" % self.ndim)
        self.in_content_original_shape = in_content.shape
        for a in range(self.ndim):
            if self.dim[a] > in_content.shape[a]:
                raise ValueError("patch larger than the volume along axis %d" % a)
        pshape = patch_array_shape(in_content.shape, self.dim, self.stride)
        self.in_content_cropped_shape = in_content_cropped_shape(in_content.shape, self.dim, self.stride)
        src_dtype = in_content.dtype
        dev = _device()
        vol = torch.from_numpy(np.ascontiguousarray(in_content, dtype=np.float64)).to(dev)
        out = torch.empty(pshape, dtype=torch.float64, device=dev)
        _lib.call("dpi_patch_extract_f64", C.c_void_p(vol.data_ptr()), _i32x3(self._pad3(in_content.shape)),
                  _i32x3(self._pad3(self.dim)), _i32x3(self._pad3(self.stride)), float(gain),
                  C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        res = out.cpu().numpy()
        if src_dtype != np.float64:
            res = res.astype(src_dtype)
        self.patch_array_shape = res.shape
        return res

    def reconstruct(self, patch_array: np.ndarray, gain: float = 1.0) -> np.ndarray:
        if not isinstance(patch_array, np.ndarray):
            raise ValueError("patch_array must be of type: " + str(np.ndarray))
        ndim = patch_array.ndim // 2
        if ndim != self.ndim:
            raise ValueError("patch_array must have %d dimensions" % (2 * self.ndim))
        idx = patch_array.shape[:ndim]
        computed = tuple(int((idx[a] - 1) * self.stride[a] + self.dim[a]) for a in range(ndim))
        if self.in_content_cropped_shape is not None and tuple(self.in_content_cropped_shape) != computed:
            raise ValueError("There is something wrong with the dimensions!")
        if tuple(patch_array.shape[ndim:]) != self.dim:
            raise ValueError("There is something wrong with the dimensions!")
        dev = _device()
        pa = torch.from_numpy(np.ascontiguousarray(patch_array, dtype=np.float32)).to(dev)
        vol = torch.empty(computed, dtype=torch.float32, device=dev)
        _lib.call("dpi_patch_reassemble_f32", C.c_void_p(pa.data_ptr()), _i32x3(self._pad3(computed)),
                  _i32x3(self._pad3(self.dim)), _i32x3(self._pad3(self.stride)), float(gain),
                  C.c_void_p(vol.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        return vol.cpu().numpy()


def _get_patch_extractor(in_shape, patch_shape, patch_stride, datadim: str, imgchannel: int = None) -> PatchExtractor:
    """``_get_patch_extractor`` (data.py:8-17): -1 -> full axis; 2.5-D: last patch axis := imgchannel."""
    ndim = len(in_shape)
    pshape = [patch_shape[d] if patch_shape[d] != -1 else in_shape[d] for d in range(ndim)]
    if datadim == "2.5d" and imgchannel is not None:
        pshape[-1] = imgchannel
    pstride = [patch_stride[d] if patch_stride[d] != -1 else pshape[d] for d in range(len(pshape))]
    return PatchExtractor(dim=tuple(pshape), stride=tuple(pstride))


def _transpose_patches_25d(a: np.ndarray, slice: str = "XY", adj: bool = False) -> np.ndarray:
    """``_transpose_patches_25d`` (data.py:20-41)"""
    s = slice.lower()
    s = {"xt": "tx", "yt": "ty"}.get(s, s)
    if s == "xy":
        return a.transpose((0, 3, 1, 2)) if adj else a.transpose((0, 2, 3, 1))
    if s == "ty":
        return a.transpose((0, 1, 3, 2))
    return a


def extract_patches(args) -> List[dict]:
    """``extract_patches`` (data.py:44-84): list of ``{'image', 'mask', 'name'}`` dicts."""
    original = np.load(os.path.join(args.imgdir, args.imgname), allow_pickle=True)
    corrupted = np.load(os.path.join(args.imgdir, args.maskname), allow_pickle=True)
    assert original.shape == corrupted.shape, "Original and Corrupted data must have the same dimension"
    assert original.ndim in [2, 3], "Data volumes have to be 2D or 3D"
    if np.isnan(corrupted).any():
        corrupted = u.bool2bin(corrupted)
    pe = _get_patch_extractor(original.shape, args.patch_shape, args.patch_stride, args.datadim, args.imgchannel)
    if args.datadim == "2.5d" or (args.datadim == "2d" and pe.ndim == 3):
        final_shape = (-1,) + pe.dim
    else:
        final_shape = (-1,) + pe.dim + (1,)
    patches_img = pe.extract(original, gain=args.gain).reshape(final_shape)
    patches_msk = pe.extract(corrupted).reshape(final_shape)
    if args.datadim == "2.5d":
        patches_img = _transpose_patches_25d(patches_img, args.slice)
        patches_msk = _transpose_patches_25d(patches_msk, args.slice)
    n = patches_img.shape[0]
    z = u.ten_digit(n)
    outputs = []
    for p in range(n):
        m = patches_msk[p]
        if args.adirandel > 0:
            m = u.add_rand_mask(m, args.adirandel)
        outputs.append({"image": patches_img[p], "mask": m, "name": str(p).zfill(z)})
    return outputs


_ALIAS_HISTORY = type("History", (u.History,), {"__module__": "utils.metrics"})


@contextlib.contextmanager
def history_alias():
    """``*_run.npy`` pickles its History by class path ``utils.metrics.History`` (main.py:226-235) so that files are
    interchangeable with the reference.  Inside this context that module path resolves — to the reference's own class
    when ``utils.metrics`` is importable, otherwise to an alias of this package's History — and ``sys.modules`` is put
    back on exit, so a caller's own ``utils`` package is never shadowed outside the pickle step.  Yields the class."""
    try:
        import utils.metrics as um
        if hasattr(um, "History"):
            yield um.History
            return
    except Exception:
        pass
    saved = {k: sys.modules.get(k) for k in ("utils", "utils.metrics")}
    pkg = types.ModuleType("utils")
    pkg.__path__ = []
    mod = types.ModuleType("utils.metrics")
    mod.History = _ALIAS_HISTORY
    pkg.metrics = mod
    sys.modules["utils"], sys.modules["utils.metrics"] = pkg, mod
    try:
        yield _ALIAS_HISTORY
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_run(path) -> dict:
    """read a ``<name>_run.npy`` result file (main.py:226-235) written by this package or by the reference: the pickled
    ``History`` resolves under ``utils.metrics`` for the duration of the load only"""
    with history_alias():
        return np.load(path, allow_pickle=True).item()


def reconstruct_patches(args, return_history: bool = False, verbose: bool = False):
    """``reconstruct_patches`` (data.py:87-130)."""
    inputs = np.load(os.path.join(args.imgdir, args.imgname), allow_pickle=True)
    pe = _get_patch_extractor(inputs.shape, args.patch_shape, args.patch_stride, args.datadim, args.imgchannel)
    pe.in_content_cropped_shape = in_content_cropped_shape(inputs.shape, pe.dim, pe.stride)
    pas = patch_array_shape(inputs.shape, pe.dim, pe.stride)
    patches_out, elapsed, history = [], [], []
    out = None
    for path in sorted(glob(os.path.join("./results", args.outdir) + "/*.npy")):
        if "output" in os.path.basename(path):
            continue
        out = load_run(path)
        o = np.asarray(out["output"], dtype=np.float32)
        patches_out.append(o)
        elapsed.append(out.get("elapsed", out.get("elapsed time")))
        history.append(out["history"])
    if not patches_out:
        raise FileNotFoundError("no *_run.npy results under ./results/%s" % args.outdir)
    # All-zero patches are written without optimising as ``img * mask`` WITH the channel axis, optimised ones without it
    # (main.py:281-284 vs 219): the reference's ``np.asarray`` of that ragged list fails.  Outputs are brought to the shape
    # of the first one that has the extractor's rank.
    shp = next((p.shape for p in patches_out if p.ndim == len(pe.dim)), None)
    shapes = sorted({p.shape for p in patches_out})
    if len(shapes) > 1 and (shp is None or any(p.size != int(np.prod(shp)) for p in patches_out)):
        raise ValueError("reconstruct_patches: the result files under ./results/%s hold outputs of different shapes %s"
                         % (args.outdir, shapes))
    patches_out = np.asarray([p.reshape(shp) if shp is not None and p.size == int(np.prod(shp)) else p
                              for p in patches_out])
    if args.datadim == "2.5d":
        patches_out = _transpose_patches_25d(patches_out, args.slice, adj=True)
    outputs = pe.reconstruct(patches_out.reshape(pas), gain=args.gain)
    if verbose:
        print("\n%d patches; total elapsed time on %s: %s" % (
            len(history), out["device"], u.sec2time(sum(u.time2sec(e) for e in elapsed))))
    return (outputs, history) if return_history else outputs
