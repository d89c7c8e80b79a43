"""Batch-sharded multi-GPU plumbing (one process per GPU, torch.distributed / NCCL).

The hot path shards by clip: rank r holds B/G clips of both modalities and the fuser
weights are replicated.  The only forward exchange is one all-reduce of a packed
float32 buffer  [sum|rgb| (C) || sum|depth| (C) || sum erank (1) || rows (1)]  -- the
global-scope channel score (the single-process reference on the concatenated batch)
and the batch-mean effective-rank statistic.  Backward: one all-reduce of the flat
fuser gradients.  ``score_scope='local'`` reproduces ``nn.DataParallel`` (each replica
scores its own shard, reference: main_utkinects.py:129) and needs no forward collective.

Everything here is backend-agnostic host logic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `total` clips for `rank`; the first total % world ranks get one extra."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_statistics(score_sums: torch.Tensor, erank_sum: torch.Tensor, rows: int) -> torch.Tensor:
    """(2, C) column sums of |x|, scalar sum of per-sample erank, local row count -> (2C + 2,) float32."""
    C = score_sums.shape[-1]
    out = torch.empty(2 * C + 2, dtype=torch.float32, device=score_sums.device)
    out[: 2 * C] = score_sums.reshape(-1)
    out[2 * C] = erank_sum
    out[2 * C + 1] = float(rows)
    return out


def unpack_statistics(packed: torch.Tensor, samples: int):
    """-> (score (2, C) = sums / global rows, mean erank over `samples` global samples)."""
    C = (packed.numel() - 2) // 2
    rows = packed[2 * C + 1]
    return packed[: 2 * C].reshape(2, C) / rows, packed[2 * C] / float(max(samples, 1))


def allreduce_statistics(packed: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


def global_bottomk_indices(score: torch.Tensor, k: int) -> torch.Tensor:
    """Every rank runs the same deterministic bottom-k on the all-reduced score, so the selected channels
    agree bit-exactly across ranks.  CUDA tensors use the sm_100a kernel; CPU tensors (the gloo tests of
    this host logic) use the equivalent stable sort -- ascending score, ties to the lower index."""
    if score.is_cuda:
        from . import ops
        return ops.bottomk(score, k)
    return torch.sort(score, dim=-1, stable=True)[1][..., :k].contiguous()


class GradBucket:
    """Flat fp32 gradient bucket for the fuser's parameters: one all-reduce per step.

    Parameters that received no gradient (the reference keeps unused `modality_token`, `projection.*`,
    `fusion_conv.*`; W_q / W_k rows of qkv get exact zeros) contribute zeros, so all ranks reduce the same
    layout without `find_unused_parameters`."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)

    def allreduce(self, average: bool = True):
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        world = 1
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.group)
            if world > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if average and world > 1:
            self.flat.div_(world)
        off = 0
        for p in self.params:
            n = p.numel()
            g = self.flat[off:off + n].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        return self.flat
