"""Import shim that lets the UNMODIFIED reference (/root/reference) run in this container.

TEST INFRASTRUCTURE.  Used only by ``oracle/gen_golden.py`` and ``tests/test_oracle_vs_reference.py``
(both skip when /root/reference is absent, e.g. on the GPU box).  It stubs the optional third-party
imports the reference pulls in at import time but the hot path never needs:
GPUtil, termcolor, skimage.util, matplotlib, imageio, and restores ``np.float``.
"""
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("DPI_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "main.py"))


class _Anything:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)

    def _ga(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything()

    m.__getattr__ = _ga
    sys.modules[name] = m
    return m


def _view_as_windows(arr, window_shape, step=1):
    from numpy.lib.stride_tricks import sliding_window_view
    if isinstance(step, int):
        step = (step,) * arr.ndim
    v = sliding_window_view(arr, tuple(window_shape))
    return v[tuple(slice(None, None, s) for s in step)]


def _view_as_blocks(arr, block_shape):
    return _view_as_windows(arr, block_shape, tuple(block_shape))


def install():
    """Make ``import main`` / ``import architectures`` / ``import utils`` resolve to the reference."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    if not hasattr(np, "float"):
        np.float = float
    if "GPUtil" not in sys.modules:
        _stub("GPUtil", getFirstAvailable=lambda *a, **k: [0], getGPUs=lambda: [])
    try:
        importlib.import_module("termcolor")
    except ImportError:
        _stub("termcolor", colored=lambda s, *a, **k: s)
    try:
        importlib.import_module("skimage.util")
    except ImportError:
        _stub("skimage")
        _stub("skimage.util", view_as_windows=_view_as_windows, view_as_blocks=_view_as_blocks)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "imageio",
                 "mpl_toolkits", "mpl_toolkits.axes_grid1"):
        try:
            importlib.import_module(name)
        except ImportError:
            _stub(name)
    if os.environ.get("CUDA_VISIBLE_DEVICES", None) == "":
        os.environ.pop("CUDA_VISIBLE_DEVICES")
    # drop file-less placeholder modules named like the reference's packages (the product registers a
    # `utils.metrics` alias for unpickling *_run.npy when the reference is not importable)
    for name in list(sys.modules):
        if name.split(".")[0] in ("utils", "architectures", "data", "parameter", "main") and \
                not getattr(sys.modules[name], "__file__", None):
            del sys.modules[name]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def reference_modules():
    """Returns (architectures, utils, data, parameter, main) of the reference."""
    install()
    import architectures  # noqa
    import utils  # noqa
    import data  # noqa
    import parameter  # noqa
    import main  # noqa
    return architectures, utils, data, parameter, main
