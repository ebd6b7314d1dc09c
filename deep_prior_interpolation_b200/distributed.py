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
    """Shared-network mode: identical weights on every rank, per-rank patch rows, one gradient all-reduce per
    iteration between backward and Adam.  The all-reduce runs on the engine's stream, so it is ordered after the
    last wgrad kernel and before the fused Adam kernel without any host synchronisation."""

    def __init__(self, engines, lr: float = 1e-3):
        # `engines`: one Engine per local patch row, all compiled from networks that share ONE FlatParams
        self.engines = list(engines)
        self.lr = lr
        self.broadcast_parameters()

    def broadcast_parameters(self, src: int = 0):
        """every rank starts from rank `src`'s weights and BatchNorm buffers (the networks are constructed per rank)"""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            P = self.engines[0].params
            dist.broadcast(P.P, src)
            dist.broadcast(P.B, src)

    def iteration(self, sigma: float):
        e0 = self.engines[0]
        P = e0.params
        acc = None
        for k, e in enumerate(self.engines):
            if sigma > 0:
                e.perturb_input(sigma)
            e.run_forward()
            e.run_loss()
            e.run_backward()
            if len(self.engines) > 1:
                acc = P.G.clone() if acc is None else acc.add_(P.G)
        if acc is not None:
            P.G.copy_(acc.div_(len(self.engines)))
        allreduce_mean_(P.G)
        e0.adam_step()
        for e in self.engines:
            e.iteration_end()
