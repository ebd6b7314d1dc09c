"""Command-line surface of the reference, kept flag-for-flag (``parameter.py:4-130``) plus a few B200 flags.

Every reference flag keeps its name, type, choices and default; the post-processing of ``parameter.py:113-125``
is reproduced.  One fix: ``--netdir`` defaults to ``[]`` instead of ``None`` so the
from-scratch path does not crash on ``len(None)``.  New flags default to reference behaviour.
"""
from __future__ import annotations

from argparse import ArgumentParser, Namespace
from typing import Optional, Sequence

_ACTS = ["LeakyReLU", "ReLU", "ELU", "Tanh", "Sigmoid"]


def build_parser() -> ArgumentParser:
    p = ArgumentParser()
    # dataset
    p.add_argument("--imgdir", type=str, required=True, default="./datasets/", help="Directory containing the processed data")
    p.add_argument("--outdir", type=str, required=False, help="Subfolder in ./results/ for saving.")
    p.add_argument("--imgname", type=str, help="The name of original images")
    p.add_argument("--maskname", type=str, help="The name of corrupted images")
    p.add_argument("--gain", type=float, required=False, default=2e3, help="gain for the input")
    p.add_argument("--datadim", type=str, required=False, default="2d", choices=["2d", "2.5d", "3d"], help="The dimensionality of the data")
    p.add_argument("--slice", type=str, required=False, default="xy", choices=["tx", "ty", "xy"], help="The type of slice of 3D data when datadim=2.5d")
    p.add_argument("--imgchannel", type=int, required=False, help="Number of 2.5d patches to be stacked in the channel dimension.")
    p.add_argument("--adirandel", type=float, required=False, default=0., help="The percent of addictive random deleting samples")
    p.add_argument("--padwidth", type=int, required=False, default=0, help="(parsed, unused — as in the reference)")
    p.add_argument("--patch_shape", nargs="+", type=int, required=False, help="Patch shape to be processed (2D, 2.5D, 3D)")
    p.add_argument("--patch_stride", nargs="+", type=int, required=False, help="Patch stride for the extraction (2D, 2.5D, 3D)")
    # network design
    p.add_argument("--net", type=str, required=False, default="multiunet", choices=["multiunet", "attmultiunet", "part", "unet", "load"], help="The network architecture")
    p.add_argument("--gpu", type=int, required=False, default=-1, help="GPU to use (-1: LOCAL_RANK / device 0)")
    p.add_argument("--activation", type=str, default="LeakyReLU", required=False, choices=_ACTS, help="Activation function of the convolution block")
    p.add_argument("--last_activation", type=str, required=False, choices=_ACTS, help="Activation function of the network output")
    p.add_argument("--dropout", type=float, default=0., required=False, help="Dropout rate (only 0 is accelerated)")
    p.add_argument("--filters", nargs="+", type=int, required=False, default=[16, 32, 64, 128, 256], help="Channels of every encoder/decoder scale")
    p.add_argument("--skip", nargs="+", type=int, required=False, default=[16, 32, 64, 128], help="Channels of the skip connections")
    p.add_argument("--inputdepth", type=int, required=False, default=64, help="Depth of the input noise tensor")
    p.add_argument("--upsample", type=str, required=False, default="nearest", choices=["nearest", "linear"], help="Decoder upsampling strategy")
    p.add_argument("--inittype", type=str, required=False, default="xavier", choices=["xavier", "normal", "default", "kaiming", "orthogonal"], help="Weight initialisation")
    p.add_argument("--initgain", type=float, required=False, default=0.02, help="Initialisation scaling factor")
    p.add_argument("--savemodel", action="store_true", default=False, help="Save the optimised model to disk")
    p.add_argument("--netdir", type=str, nargs="+", required=False, default=[], help="Path(s) for loading optimised networks")
    # input noise
    p.add_argument("--param_noise", action="store_false", help="(no-op in the reference loop; kept for compatibility)")
    p.add_argument("--reg_noise_std", type=float, required=False, default=0.03, help="Std of the noise added to the input every iteration")
    p.add_argument("--noise_dist", type=str, default="n", required=False, choices=["n", "u", "c"], help="Input noise distribution")
    p.add_argument("--noise_std", type=float, default=.1, required=False, help="Std of the input noise tensor")
    p.add_argument("--data_forgetting_factor", type=int, default=0, required=False, help="Duration of additional decimated data in the input")
    p.add_argument("--filter_noise_with_wavelet", action="store_true", default=False, help="Filter the input noise with the wavelet bandwidth")
    p.add_argument("--lowpass_fs", type=float, required=False, help="Butterworth LPF sampling frequency")
    p.add_argument("--lowpass_fc", type=float, required=False, help="Butterworth LPF cutoff frequency")
    p.add_argument("--lowpass_ntaps", type=int, required=False, default=7, help="Low pass filter length")
    # training
    p.add_argument("--loss", type=str, required=False, choices=["mae", "mse"], default="mae", help="Loss function")
    p.add_argument("--epochs", "-e", "--iter", type=int, required=False, default=2001, help="Number of optimisation iterations")
    p.add_argument("--lr", type=float, default=1e-3, required=False, help="Learning rate of Adam")
    p.add_argument("--lr_factor", type=float, default=.9, required=False, help="LR reduction for the plateau scheduler")
    p.add_argument("--lr_thresh", type=float, default=1e-5, required=False, help="LR threshold for the plateau scheduler")
    p.add_argument("--lr_patience", type=int, default=100, required=False, help="LR patience for the plateau scheduler")
    p.add_argument("--save_every", type=int, required=False, help="Save intermediate outputs every N iterations")
    p.add_argument("--start_from_prev", action="store_true", default=False, help="Start each patch from the previous patch's network")
    p.add_argument("--reduce_lr", action="store_true", default=False, help="Use the ReduceLROnPlateau scheduler")
    p.add_argument("--earlystop_patience", type=int, required=False, help="Early stopping patience")
    p.add_argument("--earlystop_min_delta", type=float, required=False, default=1., help="Early stopping min percentage delta")
    # POCS (main_pocs.py only; parsed for compatibility)
    p.add_argument("--pocs_alpha", type=float, required=False, default=0.1)
    p.add_argument("--pocs_thresh", type=float, required=False, default=5.)
    p.add_argument("--pocs_weight", type=float, required=False)
    # B200 additions (defaults reproduce the reference's behaviour)
    p.add_argument("--precision", type=str, default="fp32", choices=["fp32", "tf32"],
                   help="conv operand precision: fp32 CUDA-core kernels or tcgen05 TF32 (fp32 accumulate)")
    p.add_argument("--sync_every", type=int, default=1,
                   help="read loss/SNR/PCORR back every N iterations (1 = every iteration like the reference)")
    p.add_argument("--noise_seed", type=int, default=0, help="Philox seed of the per-iteration input noise")
    p.add_argument("--no_cuda_graph", action="store_true", default=False, help="launch kernels eagerly instead of replaying a CUDA graph")
    p.add_argument("--patches_in_flight", type=int, default=1,
                   help="independent patches optimised concurrently on one GPU, each on its own stream: 1 = one at a "
                        "time like the reference (default); 0 = choose from the patch size (3 up to 80^3 voxels, 2 up to "
                        "128^3); with K > 1 per-iteration log lines are suppressed and a patch's `elapsed` is wall time "
                        "while it shares the GPU with K-1 others")
    p.add_argument("--shared_net", action="store_true", default=False,
                   help="one network for all patches, gradients all-reduced over ranks (config 5; not in the reference)")
    return p


def finalize_args(args: Namespace) -> Namespace:
    """post-processing of parameter.py:113-125"""
    if args.upsample == "linear":
        args.upsample = "trilinear" if args.datadim == "3d" else "bilinear"
    if args.patch_shape is None:
        args.patch_shape = [-1, -1] if args.datadim == "2d" else [-1, -1, -1]
    if args.patch_stride is None:
        args.patch_stride = args.patch_shape
    if args.earlystop_patience is None:
        args.earlystop_patience = args.epochs
    if args.netdir is None:
        args.netdir = []
    return args


def parse_arguments(argv: Optional[Sequence[str]] = None) -> Namespace:
    """``parse_arguments`` (parameter.py:4-130)."""
    return finalize_args(build_parser().parse_args(argv))


_KEYS_MUST = ["datadim", "slice", "imgchannel", "patch_shape", "inputdepth", "loss", "lr", "lr_factor", "lr_thresh",
              "lr_patience", "reduce_lr"]
_KEYS_MILD = ["net", "activation", "last_activation", "dropout", "filters", "skip", "upsample", "inittype", "initgain"]


def net_args_are_same(args1: Namespace, args2: Namespace) -> bool:
    """checkpoint-compatibility check of ``parameter.py:133-173``"""
    d1, d2 = vars(args1), vars(args2)
    errors = [k for k in _KEYS_MUST if d1[k] != d2[k]]
    warnings = [k for k in _KEYS_MILD if d1[k] != d2[k]]
    if errors:
        print("The following arguments keys have to be the same:\n\t")
        print(", ".join(errors))
        return False
    if warnings:
        print("\nThe following arguments are different, but they are overridden by the network loading:")
        print("\t", ", ".join(warnings))
    return True
