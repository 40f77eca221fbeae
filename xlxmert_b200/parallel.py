"""Data-parallel gradient exchange: one process per GPU, one all-reduce per step.

The reference wraps the model in ``DistributedDataParallel(find_unused_parameters=True)``
(``x-lxmert/src/pretrain/lxmert_pretrain.py:102-106``): default 25 MB buckets discovered by an autograd-graph walk
every step.  Here the encoder's backward already writes all of its parameter gradients into ONE flat fp32 arena
(``B200LxmertEncoder.last_grad_arena``; the ``.grad`` tensors are views into it), so the bulk of the 843 MB exchange
is a single NCCL all-reduce over NVLink with no bucketing logic; the remaining parameters (embeddings, pooler, heads,
``mask_feat``) go through one more flat buffer.  Parameters without a gradient this step (the task-dependent unused
sets) are simply absent — the same "skip" semantics as ``param.grad = None`` in the reference loop
(``lxmert_pretrain.py:363-364``).

Works with any ``torch.distributed`` backend (NCCL on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def _encoders(model) -> List:
    return [m for m in model.modules() if type(m).__name__ == "B200LxmertEncoder"]


def finish_overlapped_sync(model: Optional[torch.nn.Module] = None) -> None:
    """Make the current stream wait for the stage-wise all-reduces an encoder backward left in flight
    (``enable_overlapped_gradient_sync(..., defer_wait=True)``).  Called by :func:`allreduce_gradients` and by
    ``B200AdamW.step``; call it yourself before reading encoder gradients any other way."""
    from .encoder import finish_pending_gradient_sync
    finish_pending_gradient_sync()


def allreduce_gradients(model: torch.nn.Module, group: Optional[dist.ProcessGroup] = None,
                        average: bool = True) -> int:
    """All-reduce (mean) every gradient of ``model`` across the data-parallel group.  Returns the number of
    collectives issued (2 at most: the encoder arena and the flat buffer of everything else)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    calls = 0
    nccl_avg = average and dist.get_backend(group) == "nccl"     # the mean inside the collective: no extra pass
    in_arena = set()
    for enc in _encoders(model):
        arena = enc.last_grad_arena
        if arena is None:
            continue
        live = [p for p in enc.parameters() if p.grad is not None
                and p.grad.untyped_storage().data_ptr() == arena.untyped_storage().data_ptr()]
        if not live:
            continue
        in_arena.update(id(p) for p in live)
        if getattr(enc, "arena_reduced", False):     # already exchanged stage by stage inside the backward
            continue
        dist.all_reduce(arena, group=group)
        if average:
            arena.mul_(1.0 / world)
        calls += 1
    rest = [p for p in model.parameters() if p.grad is not None and id(p) not in in_arena]
    # a tied parameter (decoder.weight ≡ word_embeddings.weight) appears once in model.parameters()
    if rest:
        flat = torch.cat([p.grad.reshape(-1) for p in rest])     # packed while the encoder's last range is still reducing
        dist.all_reduce(flat, op=dist.ReduceOp.AVG if nccl_avg else dist.ReduceOp.SUM, group=group)
        if average and not nccl_avg:
            flat.mul_(1.0 / world)
        off = 0
        for p in rest:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        calls += 1
    finish_overlapped_sync(model)
    return calls


def enable_overlapped_gradient_sync(model: torch.nn.Module, group=True, defer_wait: bool = False) -> None:
    """Make every native encoder inside ``model`` all-reduce its gradient arena during its own backward, one stage
    (cross-modality / vision / language / visual-feature layers) at a time, so the NVLink traffic hides behind the
    remaining backward compute — what DDP's bucketed hooks do for the reference (lxmert_pretrain.py:102-106).
    ``allreduce_gradients`` afterwards only handles the parameters outside the encoders.
    ``defer_wait`` (opt-in): the backward does not make the caller's stream wait for its last reduce; the wait happens
    in ``allreduce_gradients`` / ``B200AdamW.step`` (``finish_overlapped_sync``), so the tail of the backward and the
    packing of the remaining gradients overlap it.  Covered by a single-GPU test of the bookkeeping
    (``tests/test_encoder_parity.py``); not yet measured on several GPUs, hence off by default."""
    for enc in _encoders(model):
        enc.grad_sync_group = group
        enc.defer_sync_wait = bool(defer_wait)


def shard_batch(batch: dict, rank: int, world: int) -> dict:
    """Rank ``rank``'s disjoint slice of a batch dict (what ``DistributedSampler`` does to the dataset,
    ``lxmert_data.py:663-667``): contiguous equal shards, remainder dropped."""
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v):
            n = v.shape[0] // world
            out[k] = v[rank * n:(rank + 1) * n]
        else:
            out[k] = v
    return out
