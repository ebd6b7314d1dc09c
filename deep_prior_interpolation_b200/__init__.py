"""deep_prior_interpolation_b200 — B200-native hot path of polimi-ispl/deep_prior_interpolation.

Host side (Python) mirrors the reference's surface for the path: ``parse_arguments``, ``get_net`` /
``MulResUnet3D`` / ``MulResUnet``, ``Interpolator``, ``extract_patches`` / ``reconstruct_patches``.
Compute is the C-ABI library ``_C/libdpi_b200.so`` (``include/dpi_b200.h``): hand-written sm_100a kernels.
Importing the package fails loudly if that library has not been built — there is no fallback.
"""
from . import _lib  # noqa: F401  (loads libdpi_b200.so or raises ImportError)
from .architectures import get_net, MulResUnet, MulResUnet3D, AttMulResUnet2D, DeepPriorNet  # noqa: F401

__version__ = "0.1.0"
