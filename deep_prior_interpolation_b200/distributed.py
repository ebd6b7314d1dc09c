"""One-process-per-GPU plumbing (``torch.distributed``; NCCL over NVLink on the B200 box, gloo in CPU tests).

Two modes:

* **patch sharding** — the reference's patches are independent units (own network, own noise, own Adam state;
  ``main.py:274-295``), so rank ``r`` of ``W`` takes the patches ``p % W == r`` and there is NO data-path
  collective; results meet only as ``<name>_run.npy`` files.
* **shared network** (BASELINE config 5; not in the reference) — one network, batch rows = patches, every rank
  runs forward/backward on its rows, the flat 5.9 M-float gradient buffer is averaged with ONE all-reduce per
  iteration, then every rank applies the identical fused Adam step.  BatchNorm statistics stay local to the rank
  ("local-BN", the DDP default); see ``SharedNetTrainer``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import torch
import torch.distributed as dist


def rank_world() -> tuple:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def patch_indices(n_patches: int, rank: Optional[int] = None, world: Optional[int] = None) -> List[int]:
    """round-robin assignment of patch indices to ranks"""
    if rank is None or world is None:
        rank, world = rank_world()
    return list(range(rank, n_patches, world))


def init_process_group(device: Optional[torch.device] = None, backend: Optional[str] = None):
    if dist.is_initialized():
        return
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend == "nccl" and device is not None:
        dist.init_process_group(backend, device_id=device)
    else:
        dist.init_process_group(backend)


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """in-place average over ranks (gradient of the global-mean loss when every rank holds equally many rows)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    return flat


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class SharedNetTrainer:
    """Shared-network mode (BASELINE config 5; no reference equivalent - it replaces the per-patch networks of
    ``main.py:274-295`` by ONE network optimised over all patches): batch rows = patches, rank ``r`` holds the rows
    ``p % W == r``.  One iteration:

      graph A   for every local row, in row order: perturb its z, forward, masked loss, backward (``Engine.run_row``),
                then ``acc (+)= g_row / N`` with N = the GLOBAL number of rows (``dpi_axpby``; fixed order, no atomics)
      NCCL      ONE all-reduce(sum) of the flat 5.9 M-float buffer -> the gradient of the global-mean loss, bit-identical
                on every rank
      graph B   the fused Adam step on the flat buffers (identical on every rank), per-row bookkeeping

    A and B are CUDA graphs; the all-reduce is issued between the two replays on the same stream, so the three are
    stream-ordered without any host synchronisation.  With ``DPI_SHARED_NET_ONE_GRAPH=1`` the collective is captured
    into a single graph together with A and B (NCCL supports capture; kept opt-in).

    BatchNorm uses the statistics of the row it is normalising ("local-BN", per row: every row is a forward pass of its
    own), so the result equals, for every row, a forward/backward of the reference network fed that row alone, with the
    gradients averaged over all rows (the emulation of tests/test_gpu_shared_net.py).  Running statistics are local to
    a rank (they are not used by training-mode BatchNorm); rank 0's are the ones saved with the model."""

    def __init__(self, engine, rows, n_rows_global: int, lr: float = 1e-3, sigma: float = 0.03, group=None):
        self.eng, self.rows, self.N = engine, list(rows), int(n_rows_global)
        self.lr, self.sigma, self.group = float(lr), float(sigma), group
        P = engine.params
        # (a rank without rows still takes part in the collective with a zero contribution)
        self.acc = torch.zeros_like(P.G)
        self.graph_a = self.graph_b = self.graph_all = None
        self.one_graph = os.environ.get("DPI_SHARED_NET_ONE_GRAPH", "0") == "1"
        self.broadcast_parameters()

    @property
    def distributed(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def broadcast_parameters(self, src: int = 0):
        """every rank starts from rank `src`'s weights and BatchNorm buffers (the networks are constructed per rank)"""
        if self.distributed:
            P = self.eng.params
            dist.broadcast(P.P, src, group=self.group)
            dist.broadcast(P.B, src, group=self.group)

    def reset(self):
        self.eng.reset_loop_state(self.lr, self.rows[0].seed if self.rows else 0)
        for r in self.rows:
            r.reset()

    # ---- the three stages ----------------------------------------------------------------------------------
    def _stage_a(self, st=None):
        from . import _lib
        eng, P = self.eng, self.eng.params
        stp = C.c_void_p(eng.stream if st is None else st)
        if not self.rows:
            self.acc.zero_()
        for k, r in enumerate(self.rows):
            eng.run_row(r, self.sigma, st)
            _lib.call("dpi_axpby", C.c_void_p(self.acc.data_ptr()), C.c_void_p(P.G.data_ptr()), 0.0 if k == 0 else 1.0,
                      1.0 / self.N, P.n, stp)

    def _allreduce(self):
        if self.distributed:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM, group=self.group)

    def _stage_b(self, st=None):
        from . import _lib
        eng, P = self.eng, self.eng.params
        _lib.call("dpi_adam_step_dev", C.c_void_p(P.P.data_ptr()), C.c_void_p(self.acc.data_ptr()),
                  C.c_void_p(eng.adam_m.data_ptr()), C.c_void_p(eng.adam_v.data_ptr()), P.n,
                  C.c_void_p(eng.hyper.data_ptr()), 0.9, 0.999, 1e-8, 0.0, C.c_void_p(eng.stream if st is None else st))
        if not self.rows:
            # keep the Adam step / learning-rate cell of a row-less rank in step with the others
            eng.iteration_end(st)
        for r in self.rows:
            eng.row_end(r, st)

    # ---- execution -----------------------------------------------------------------------------------------
    def iteration_eager(self):
        self._stage_a()
        self._allreduce()
        self._stage_b()

    def capture(self):
        eng = self.eng
        eng._refresh()
        if self.distributed:
            # the communicator must exist before anything is captured
            dist.all_reduce(torch.zeros(4, device=eng.device), group=self.group)
        torch.cuda.synchronize(eng.device)
        s = torch.cuda.Stream(eng.device)
        s.wait_stream(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(s):
            if self.one_graph:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    cs = torch.cuda.current_stream(eng.device).cuda_stream
                    self._stage_a(cs)
                    self._allreduce()
                    self._stage_b(cs)
                self.graph_all = g
            else:
                ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(ga, stream=s):
                    self._stage_a(torch.cuda.current_stream(eng.device).cuda_stream)
                with torch.cuda.graph(gb, stream=s):
                    self._stage_b(torch.cuda.current_stream(eng.device).cuda_stream)
                self.graph_a, self.graph_b = ga, gb
        torch.cuda.current_stream(eng.device).wait_stream(s)

    def iteration(self):
        """one optimisation iteration over all rows of all ranks (graphs must have been captured)"""
        if self.graph_all is not None:
            self.graph_all.replay()
            return
        self.graph_a.replay()
        self._allreduce()
        self.graph_b.replay()

    def param_checksum(self):
        """(sum, sum of squares) of the flat parameter buffer in float64: equal on every rank iff the weights are"""
        p = self.eng.params.P.double()
        return float(p.sum()), float((p * p).sum())
