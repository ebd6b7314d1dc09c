"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, flags / constructors /
file formats match the reference's golden records, layout bookkeeping, patch sharding over ranks (gloo)."""
import ctypes
import json
import os
import pickle
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "dpi_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(dpi_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    lib = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), "libdpi_b200.so does not export %s" % name
    from deep_prior_interpolation_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared, "ctypes signature table out of sync with include/dpi_b200.h"
    assert _lib.lib.dpi_version() >= 100
    # arity of every prototype in the header equals the ctypes table
    for name in declared:
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, hdr, re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("void", "") else len([p for p in params.split(",") if p.strip()])
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every struct of include/dpi_b200.h, as gcc lays them out, equal the ctypes mirrors in _lib.py
    (the structs cross the boundary by pointer: a drifted field would silently corrupt a launch)"""
    from deep_prior_interpolation_b200 import _lib
    mirrors = {"dpi_conv_geom": _lib.ConvGeom, "dpi_parts": _lib.Parts, "dpi_bn_next_reduce": _lib.NextReduce,
               "dpi_stats_parts": _lib.StatsParts, "dpi_pack_job": _lib.PackJob}
    hdr = open(os.path.join(ROOT, "include", "dpi_b200.h")).read()
    declared = set(re.findall(r"^}\s*(dpi_[a-z0-9_]+)\s*;", hdr, re.M))
    assert declared == set(mirrors), "a struct of the header has no ctypes mirror (or the reverse): %s" % (declared ^ set(mirrors))
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "dpi_b200.h"', "int main(void) {"]
    for cname, cls in mirrors.items():
        lines.append('  printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf(" %s=%%zu", offsetof(%s, %s));' % (fname, cname, fname))
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    assert len(out) == len(mirrors)
    for line in out:
        tok = line.split()
        cls = mirrors[tok[0]]
        assert int(tok[1]) == ctypes.sizeof(cls), (tok[0], tok[1], ctypes.sizeof(cls))
        for kv in tok[2:]:
            k, v = kv.split("=")
            assert int(v) == getattr(cls, k).offset, (tok[0], k, v, getattr(cls, k).offset)


def test_no_cpu_fallback_in_product_path():
    import deep_prior_interpolation_b200 as dpi
    from argparse import Namespace
    args = Namespace(datadim="3d", net="multiunet", upsample="trilinear", activation="LeakyReLU", last_activation=None,
                     dropout=0., inputdepth=8, filters=[4, 8, 16, 32, 64], skip=[4, 8, 16, 32])
    net = dpi.get_net(args, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 8, 32, 16, 16))
    # nothing under the product package imports the oracle
    pkg = os.path.join(ROOT, "deep_prior_interpolation_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


def test_state_dict_inventory_matches_reference():
    import deep_prior_interpolation_b200 as dpi
    from argparse import Namespace
    inv = json.load(open(os.path.join(GOLD, "state_dict_inventory.json")))
    for name, datadim, up in (("3d", "3d", "trilinear"), ("2d", "2d", "bilinear")):
        args = Namespace(datadim=datadim, net="multiunet", upsample=up, activation="LeakyReLU", last_activation=None,
                         dropout=0., inputdepth=64, filters=[16, 32, 64, 128, 256], skip=[16, 32, 64, 128])
        net = dpi.get_net(args, 1)
        mine = [[k, list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()]
        assert mine == inv[name]
        assert sum(p.numel() for p in net.parameters()) == inv[name + "_num_params"]
    assert inv["3d_num_params"] == 5923614 and inv["2d_num_params"] == 2186704


def test_attention_variant_constructor_surface():
    """--net attmultiunet (architectures/__init__.py:21-31, attention.py:197-262): the reference's key space (pinned by
    tests/golden/attnet2d_*.npz, whose `keys` are the reference's own state_dict keys), 2-D only, no CPU path"""
    import deep_prior_interpolation_b200 as dpi
    from argparse import Namespace
    g = np.load(os.path.join(GOLD, "attnet2d_full_scalars.npz"), allow_pickle=False)
    args = Namespace(datadim="2d", net="attmultiunet", upsample="nearest", activation="LeakyReLU", last_activation="Tanh",
                     dropout=0., inputdepth=64, filters=[16, 32, 64, 128, 256], skip=[16, 32, 64, 128])
    net = dpi.get_net(args, 1)
    assert list(net.state_dict().keys()) == [str(k) for k in g["keys"]]
    assert net.spec["kind"] == "attmultiunet" and sum(p.numel() for p in net.parameters()) == 2678314
    assert [n for n, _ in net.named_children()][:8] == ["down_mb1", "down_mb2", "down_mb3", "down_mb4", "down_mb5",
                                                         "down1", "up_mb1", "att1"]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 64, 32, 32))
    # with --datadim 3d the reference's get_net falls through to MulResUnet3D
    args.datadim, args.upsample, args.last_activation = "3d", "trilinear", None
    assert dpi.get_net(args, 1).spec.get("kind") is None and "4.0.weight" in dpi.get_net(args, 1).state_dict()
    args.net = "unet"
    with pytest.raises(NotImplementedError):
        dpi.get_net(args, 1)


def test_engine_plans_assemble_on_the_host(monkeypatch):
    """The launch plan of every supported architecture is built on the HOST from the module tree (no kernel runs while
    it is built), so its structure can be checked without a GPU: one ConvOp per convolution of the reference's
    state_dict (49 for MulResUnet3D), one BatchNorm finalize per BatchNorm module in each
    direction, every conv in the batched pack table, and the first-writer / accumulate resolution of the gradient
    buffers completes (it asserts on partially initialised slices)."""
    import contextlib
    from argparse import Namespace
    import deep_prior_interpolation_b200 as dpi
    from deep_prior_interpolation_b200 import engine as E

    class _NoStream:
        def __init__(self, *a, **k):
            pass

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "Stream", _NoStream)
    small = dict(inputdepth=8, filters=[4, 8, 16, 32, 64], skip=[4, 8, 16, 32])
    cases = [("multiunet", "3d", (32, 16, 16), "trilinear"), ("multiunet", "2d", (43, 25), "bilinear"),
             ("attmultiunet", "2d", (48, 32), "bilinear")]
    for kind, datadim, dims, up in cases:
        for prec in ("fp32", "tf32"):
            a = Namespace(datadim=datadim, net=kind, upsample=up, activation="LeakyReLU", last_activation=None, dropout=0.,
                          precision=prec, **small)
            net = dpi.get_net(a, 1)
            eng = E.Engine(net, dims, "cpu", precision=prec, max_iters=4)
            convs = [op for op in eng.ops if isinstance(op, E.ConvOp)]
            n_conv_w = sum(1 for k, v in net.state_dict().items() if k.endswith("weight") and v.dim() >= 4)
            n_bn = sum(1 for k in net.state_dict() if k.endswith("running_mean"))
            assert len(convs) == n_conv_w, (kind, datadim, len(convs), n_conv_w)
            if kind == "multiunet" and datadim == "3d":
                assert len(convs) == 49

            def names(calls):
                out = []
                for c in calls:
                    for cc in (c.calls if isinstance(c, E._SideCall) else [c]):
                        if isinstance(cc, E._Call):
                            out.append(cc.name)
                return out

            f, b = names(eng.fwd_calls), names(eng.bwd_calls)
            assert f.count("dpi_bn_finalize") + f.count("dpi_bn_finalize_parts") == n_bn, (kind, datadim)
            assert b.count("dpi_bn_bwd_finalize") == n_bn, (kind, datadim)
            # the BatchNorm over a block's concatenated branches (nine MultiRes blocks of the 3-D net; the 2-D blocks have
            # none, mulresunet.py:22-35): a statistics pass over the concatenation (default), or - DPI_FUSE_PART_STATS=1 -
            # the statistics its three producers left (dpi_bn_finalize_parts)
            assert f.count("dpi_channel_stats_parts") + f.count("dpi_bn_finalize_parts") == (9 if datadim == "3d" else 0)
            # every BatchNorm-backward reduce is either a launch of its own or fused into the apply pass before it
            # (dpi_bn_bwd_apply_next of kind 3 carries no reduce: it is the BatchNorm apply that also writes the gradients of
            #  both addends of a ResPath add, whose own backward launches disappear)
            kind3 = sum(1 for op in eng.ops if isinstance(op, E.AddActOp) and op.bwd_fused)
            assert kind3 == (4 if kind == "multiunet" else 0)      # one per ResPath
            fused = b.count("dpi_bn_bwd_apply_next") - kind3 + b.count("dpi_bn_bwd_apply_parts_next")
            assert b.count("dpi_bn_bwd_reduce") + b.count("dpi_bn_bwd_reduce_parts") + fused == n_bn, (kind, datadim)
            assert f.count("dpi_conv_fwd") + f.count("dpi_conv_fwd_stats") == len(convs)
            assert b.count("dpi_conv_wgrad") == len(convs) and b[-1] == "dpi_unpack_conv_wgrad_batched"
            # the two convs fed by the noise input need no data gradient
            assert b.count("dpi_conv_dgrad") == len(convs) - 2
            gates = [op for op in eng.ops if isinstance(op, E.GateMulOp)]
            assert len(gates) == (4 if kind == "attmultiunet" else 0)
            assert f.count("dpi_gate_mul_fwd") == b.count("dpi_gate_mul_bwd") == len(gates)
            if prec == "tf32":
                assert "dpi_conv_fwd_stats" in f          # BatchNorm statistics requested from the conv epilogues
    a = Namespace(datadim="2d", net="attmultiunet", upsample="bilinear", activation="LeakyReLU", last_activation=None,
                  dropout=0., precision="fp32", **small)
    with pytest.raises(ValueError, match="divisible by 16"):
        E.Engine(dpi.get_net(a, 1), (40, 24), "cpu", precision="fp32", max_iters=4)
    # four scales: sizes divisible by 8 are enough
    a.filters, a.skip = [4, 8, 16, 32], [4, 8, 16]
    eng = E.Engine(dpi.get_net(a, 1), (40, 24), "cpu", precision="fp32", max_iters=4)
    assert sum(isinstance(op, E.GateMulOp) for op in eng.ops) == 3


def test_optimize_host_logic_with_a_stub_engine(tmp_path):
    """The host side of ``Interpolator.optimize`` (main.py:195-220): chunking of the launch loop (``--sync_every``,
    saved iterations), ReduceLROnPlateau writing the new rate back to the device (``main.py:201-208,214-215``) and
    EarlyStopping (``main.py:216-217``), driven by a stub engine that serves a prescribed loss curve - no GPU."""
    from deep_prior_interpolation_b200 import interpolator as I, utils as u
    from deep_prior_interpolation_b200.parameter import parse_arguments

    class StubGraph:
        def __init__(self, eng):
            self.eng = eng

        def replay(self):
            self.eng.replays += 1

    class StubEngine:
        def __init__(self, losses):
            n = len(losses)
            self.history = torch.zeros((n, 4), dtype=torch.float64)
            self.history[:, 0] = torch.tensor(losses, dtype=torch.float64)
            self.history[:, 3] = 1e-3
            self.replays, self.lrs, self.graph = 0, [], StubGraph(self)

        def set_lr(self, lr):
            self.lrs.append(lr)
            self.history[self.replays:, 3] = lr       # the device logs its current rate with every later iteration

        def output_nchw(self, best=False):
            return torch.zeros(1, 1, 4, 4, 4)

    def make(argv, losses):
        a = parse_arguments(["--imgdir", "X"] + argv)
        T = I.Interpolator.__new__(I.Interpolator)
        T.args, T.outpath, T.image_name, T.zfill = a, str(tmp_path), "0", u.ten_digit(a.epochs)
        T.history, T.iiter, T.loss_min, T.input_list = u.History(a.epochs), 0, None, []
        T.iter_to_be_saved = list(range(0, a.epochs, int(a.save_every))) if a.save_every is not None else [0]
        p = torch.nn.Parameter(torch.zeros(1))
        T.optimizer = torch.optim.SGD([p], lr=a.lr)
        eng = StubEngine(losses)
        sched = torch.optim.lr_scheduler.ReduceLROnPlateau(T.optimizer, mode="min", factor=a.lr_factor,
                                                           threshold=a.lr_thresh, patience=a.lr_patience)
        stop = u.EarlyStopping(patience=a.earlystop_patience, min_delta=a.earlystop_min_delta, percentage=True)
        sync = max(1, int(a.sync_every))
        if a.reduce_lr or a.earlystop_patience < a.epochs:
            sync = 1
        T._opt = {"eng": eng, "scheduler": sched, "stopper": stop, "sigma": 0.03, "use_graph": True, "sync_every": sync,
                  "save_at": sorted(i for i in T.iter_to_be_saved if i != 0), "j": 0, "n": 0, "quiet": True}
        return T, eng

    def drive(T):
        chunks, done = [], False
        while not done:
            T._opt_launch()
            chunks.append(T._opt["n"])
            done = T._opt_collect()
        return chunks

    # plain run, read-back every 4 iterations: 10 iterations in chunks of 4, 4, 2; history complete
    T, eng = make(["--epochs", "10", "--sync_every", "4"], [1.0 / (i + 1) for i in range(10)])
    assert drive(T) == [4, 4, 2] and eng.replays == 10 and T.iiter == 10
    assert T.history.loss == [1.0 / (i + 1) for i in range(10)] and len(T.history.lr) == 10 and T.loss_min == 0.1

    # a plateau: the scheduler cuts the rate after `lr_patience` bad iterations and the new rate goes to the device
    curve = [1.0, 0.5] + [0.5] * 8
    T, eng = make(["--epochs", "10", "--reduce_lr", "--lr_patience", "2", "--lr_factor", "0.5", "--sync_every", "5"], curve)
    assert drive(T) == [1] * 10                       # host decisions every iteration force per-iteration read-backs
    ref_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=T.args.lr)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(ref_opt, mode="min", factor=0.5, threshold=T.args.lr_thresh, patience=2)
    want = []
    for l in curve:
        before = ref_opt.param_groups[0]["lr"]
        ref.step(l)
        if ref_opt.param_groups[0]["lr"] != before:
            want.append(ref_opt.param_groups[0]["lr"])
    assert eng.lrs == want and len(want) >= 2 and want[0] == T.args.lr * 0.5
    assert T.history.lr[0] == T.args.lr and T.history.lr[-1] in want

    # early stopping: no improvement by min_delta percent for `patience` iterations ends the loop early
    curve = [1.0, 0.9, 0.8] + [0.8] * 20
    T, eng = make(["--epochs", "23", "--earlystop_patience", "3", "--earlystop_min_delta", "1"], curve)
    ref_stop = u.EarlyStopping(patience=3, min_delta=1., percentage=True)
    n_ref = next(i + 1 for i, l in enumerate(curve) if ref_stop.step(torch.tensor(l)))
    chunks = drive(T)
    assert T.iiter == n_ref < 23 and eng.replays == n_ref and chunks == [1] * n_ref

    # saved iterations split the chunks so that the output on the device is the one to be saved
    T, eng = make(["--epochs", "12", "--save_every", "5", "--sync_every", "8"], [1.0] * 12)
    assert drive(T) == [6, 5, 1]                      # ... iteration 5 ends a chunk, then 10
    assert sorted(f for f in os.listdir(tmp_path) if "_output" in f) == ["0_output%s.npy" % str(i).zfill(T.zfill) for i in (5, 10)]


def test_patches_in_flight_policy():
    """how many independent patches share one GPU (interpolator.patches_in_flight; measured defaults)"""
    from deep_prior_interpolation_b200.interpolator import patches_in_flight
    from deep_prior_interpolation_b200.parameter import parse_arguments
    a = parse_arguments(["--imgdir", "X"])
    assert patches_in_flight(a, (64, 64, 64), 1470) == 1               # default: one at a time, like the reference
    a = parse_arguments(["--imgdir", "X", "--patches_in_flight", "0"])  # 0 = automatic
    assert patches_in_flight(a, (64, 64, 64), 1470) == 3
    assert patches_in_flight(a, (128, 128, 128), 240) == 2
    assert patches_in_flight(a, (256, 128, 128), 3) == 1
    assert patches_in_flight(a, (64, 64, 64), 2) == 2                  # never more than there are patches
    a.patches_in_flight = 5
    assert patches_in_flight(a, (256, 128, 128), 8) == 5               # an explicit flag wins
    a.start_from_prev = True
    assert patches_in_flight(a, (64, 64, 64), 1470) == 1               # sequential by definition (main.py:286)


def test_parse_arguments_matches_reference_defaults():
    from deep_prior_interpolation_b200.parameter import parse_arguments, net_args_are_same
    gold = json.load(open(os.path.join(GOLD, "parse_arguments.json")))
    new_flags = {"precision", "sync_every", "noise_seed", "no_cuda_graph", "shared_net", "patches_in_flight"}
    mine = vars(parse_arguments(["--imgdir", "X", "--netdir", "a/b"]))
    for k, v in gold["defaults"].items():
        if k == "gpu":
            assert mine[k] == -1          # documented deviation: there is no CPU path, default = auto-select
            continue
        assert mine[k] == v, (k, mine[k], v)
    assert set(mine) - set(gold["defaults"]) == new_flags
    mine2 = vars(parse_arguments(["--imgdir", "X", "--netdir", "a/b", "--datadim", "3d", "--upsample", "linear",
                                  "--patch_shape", "64", "64", "64", "--epochs", "77", "--param_noise", "--savemodel"]))
    for k, v in gold["case3d"].items():
        if k != "gpu":
            assert mine2[k] == v, (k, mine2[k], v)
    # no --netdir must not crash
    a = parse_arguments(["--imgdir", "X"])
    assert a.netdir == [] and a.patch_shape == [-1, -1] and a.earlystop_patience == a.epochs
    b = parse_arguments(["--imgdir", "X", "--lr", "0.01"])
    assert not net_args_are_same(a, b)
    c = parse_arguments(["--imgdir", "X", "--activation", "ReLU"])
    assert net_args_are_same(a, c)


def test_host_helpers_match_reference():
    from deep_prior_interpolation_b200 import utils as u
    h = json.load(open(os.path.join(GOLD, "host_helpers.json")))
    es = u.EarlyStopping(patience=3, min_delta=1., percentage=True)
    assert [bool(es.step(v)) for v in h["early_stop_seq"]] == h["early_stop"]
    assert u.EarlyStopping(patience=5).step(float("nan")) is False       # first value only sets `best`
    es2 = u.EarlyStopping(patience=5)
    es2.step(1.0)
    assert es2.step(float("nan")) is True
    for n, d in h["ten_digit"]:
        assert u.ten_digit(n) == d
    for s, t in h["sec2time"]:
        assert u.sec2time(s) == t
    hist = u.History(3000)
    hist.append((0.5, 1.25, 0.75))
    hist.lr.append(1e-3)
    assert hist.log_message(0) == h["history_msg"] and len(hist) == 1


def test_run_file_history_pickles_under_reference_class_path(tmp_path):
    from deep_prior_interpolation_b200 import utils as u
    import sys
    from deep_prior_interpolation_b200.data import history_alias
    before = {k: sys.modules.get(k) for k in ("utils", "utils.metrics")}
    with history_alias() as History:
        h = History.__new__(History)
        h.__dict__.update(u.History(10).__dict__)
        blob = pickle.dumps(h)
        assert b"utils.metrics" in blob and b"History" in blob
        np.save(tmp_path / "0_run.npy", {"history": h, "output": np.zeros((2, 2), np.float32)})
        back = np.load(tmp_path / "0_run.npy", allow_pickle=True).item()
    assert back["history"].zfill == 2
    # the alias is scoped to the pickle step: sys.modules is as it was (a caller's own `utils` is never shadowed)
    assert {k: sys.modules.get(k) for k in ("utils", "utils.metrics")} == before


def test_channel_layouts():
    from deep_prior_interpolation_b200.layout import ChannelLayout, pad4
    parts = [ChannelLayout.dense(4), ChannelLayout.dense(8), ChannelLayout.dense(13)]
    lay = ChannelLayout.concat(parts)
    assert (lay.C_l, lay.C_p) == (25, 28) and lay.part_offsets(parts) == [0, 4, 12]
    m = lay.phys2log()
    assert list(m[:12]) == list(range(12)) and list(m[12:25]) == list(range(12, 25)) and list(m[25:]) == [-1] * 3
    lay2 = ChannelLayout.concat([ChannelLayout.dense(1), ChannelLayout.dense(2), ChannelLayout.dense(3)])
    assert list(lay2.phys2log()) == [0, -1, -1, -1, 1, 2, -1, -1, 3, 4, 5, -1]
    cat = ChannelLayout.concat([ChannelLayout.dense(16), lay])
    assert cat.C_l == 41 and cat.C_p == 44 and cat.phys2log()[16] == 16 and cat.phys2log()[43] == -1
    assert [pad4(c) for c in (1, 4, 5, 25, 426, 554)] == [4, 4, 8, 28, 428, 556]
    assert hash(ChannelLayout.dense(8)) == hash(ChannelLayout.dense(8))


def test_patch_grid_helpers():
    from deep_prior_interpolation_b200 import data as d
    assert d.patch_array_shape((1000, 256, 256), (64, 64, 64), (32, 32, 32)) == (30, 7, 7, 64, 64, 64)
    assert d.count_patches((1000, 256, 256), (64, 64, 64), (32, 32, 32)) == 1470
    assert d.in_content_cropped_shape((1000, 256, 256), (64, 64, 64), (32, 32, 32)) == (992, 256, 256)
    assert d.count_patches((2000, 512, 512), (128, 128, 128), (128, 128, 128)) == 240
    pe = d._get_patch_extractor((170, 100, 1), [-1, -1, -1], [-1, -1, -1], "2.5d", 1)
    assert pe.dim == (170, 100, 1) and pe.stride == (170, 100, 1)
    a = np.arange(2 * 3 * 4 * 5).reshape(2, 3, 4, 5)
    for s in ("xy", "ty", "tx"):
        assert np.array_equal(d._transpose_patches_25d(d._transpose_patches_25d(a, s), s, adj=True), a)
    with pytest.raises(RuntimeError, match="CUDA"):
        d.PatchExtractor((2, 2)).extract(np.zeros((4, 4)))      # no silent CPU path


def test_mask_helpers():
    from deep_prior_interpolation_b200 import utils as u
    np.random.seed(4)
    data = np.random.randn(10, 8, 6)
    m = u.build_mask(data, 0.7)
    assert m.shape == data.shape and set(np.unique(m)) == {0.0, 1.0}
    assert int((m[0] == 0).sum()) == int(48 * 0.7) and np.all(m == m[0:1])       # whole traces removed
    nanv = data.copy()
    nanv[:, m[0] == 0] = np.nan
    assert np.array_equal(u.bool2bin(nanv), m)
    m2 = u.add_rand_mask(m, 0.5)
    assert m2.sum() < m.sum() and np.all(m2 <= m)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from deep_prior_interpolation_b200 import distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = D.patch_indices(11, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # shared-network mode: the flat gradient is averaged over ranks
    g = torch.arange(8, dtype=torch.float32) * (rank + 1)
    D.allreduce_mean_(g)
    stats = torch.tensor([float(rank + 1), 2.0 * (rank + 1)], dtype=torch.float64)
    D.allreduce_sum_(stats)
    q.put((rank, mine, gathered, g.tolist(), stats.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_patch_sharding_and_gradient_allreduce_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mine0, all0, g0, s0), (r1, mine1, all1, g1, s1) = res
    assert mine0 == [0, 2, 4, 6, 8, 10] and mine1 == [1, 3, 5, 7, 9]
    assert sorted(all0[0] + all0[1]) == list(range(11))
    assert g0 == g1 == [1.5 * i for i in range(8)]
    assert s0 == s1 == [3.0, 6.0]
