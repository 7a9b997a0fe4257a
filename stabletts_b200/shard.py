"""Batch sharding of the CFM solve across the GPUs of one box.

Utterances are independent on this path (LayerNorm is per frame, attention per utterance, a CFG
pair shares a sample), so the multi-GPU form is pure data parallelism: one process per GPU, weights
replicated, contiguous (or cost-balanced) batch slices, and NO collective inside the ODE loop.  The
only communication is the scatter of ``(mu, mask, c, z)`` from the root and the gather of the mel
back — a few MB over NVLink/NVSwitch through ``torch.distributed`` (backend ``nccl`` on GPUs;
``gloo`` in the CPU tests, which exercise this plumbing with a stand-in solve).
The reference has no multi-GPU inference at all; this module is new surface, not a replacement.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def utterance_cost(length: int) -> float:
    """Per-utterance estimator work in FLOPs per evaluation, BASELINE.md §3:
    F_call(T)·T = 32.948e6·T + 6144·T² (attention is quadratic in the utterance's own length)."""
    return 32.948e6 * length + 6144.0 * length * length


def split_counts(n_items: int, world: int) -> List[int]:
    """Contiguous near-equal slice sizes (first ``n_items % world`` ranks get one more)."""
    base, rem = divmod(n_items, world)
    return [base + (1 if r < rem else 0) for r in range(world)]


def partition_by_cost(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-processing-time partition of utterance indices into ``world`` shards balancing
    sum(utterance_cost); each shard's indices are returned sorted by length (bucket-friendly)."""
    order = sorted(range(len(lengths)), key=lambda i: -lengths[i])
    loads = [0.0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], len(shards[k])))
        shards[r].append(i)
        loads[r] += utterance_cost(int(lengths[i]))
    return [sorted(s, key=lambda i: lengths[i]) for s in shards]


def length_buckets(lengths: Sequence[int], n_buckets: int) -> List[List[int]]:
    """Bucket utterance indices by length (the idea of the reference's DistributedBucketSampler,
    datas/sampler.py:4-132, applied to inference): sort by length, cut into ``n_buckets`` contiguous
    groups of near-equal COST so each group pads only to its own maximum."""
    order = sorted(range(len(lengths)), key=lambda i: lengths[i])
    total = sum(utterance_cost(int(lengths[i])) for i in order)
    buckets: List[List[int]] = [[]]
    acc = 0.0
    for i in order:
        c = utterance_cost(int(lengths[i]))
        if buckets[-1] and acc + c / 2 > total * len(buckets) / n_buckets and len(buckets) < n_buckets:
            buckets.append([])
        buckets[-1].append(i)
        acc += c
    return [b for b in buckets if b]


def bucketed_solve(solve: Callable[..., torch.Tensor], mu, mask, c, z, lengths: Sequence[int], n_buckets: int = 4,
                   min_pad: int = 4) -> torch.Tensor:
    """Run ``solve`` per length bucket with each bucket cropped to (its max length + ``min_pad``) frames
    (capped at T), and scatter the results back into a (B, M, T) tensor.  Padded frames keep the initial noise ``z``
    exactly as in the reference (``z`` is unmasked, models/flow_matching.py:45, and the vector field is exactly 0 at
    masked frames, so the ODE state never moves there) — inside and beyond the crop alike.

    ``min_pad``: how many pad frames an utterance needs behind it to be insensitive to further padding.  Without
    CFG 3 is exact (cond_proj's three k=3 convs see zeros, the unmasked in_proj / long-skip convs
    reach one frame further).  With CFG the unconditional branch broadcasts a NON-ZERO ``fake_content`` over the pad
    frames too, so the last valid frame sees cond_proj at frame L, which depends on ``mu[L+3]`` — a crop to L+3
    frames replaces that by the conv's zero padding (3e-5 max-rel on valid frames, measured with the oracle); 4 pad
    frames are exact for both.  The default is therefore 4."""
    B, M, T = mu.shape
    out = z.clone()
    for idx in length_buckets(lengths, n_buckets):
        Tb = min(T, max(int(lengths[i]) for i in idx) + min_pad)
        sel = torch.as_tensor(idx, device=mu.device)
        o = solve(mu[sel, :, :Tb].contiguous(), mask[sel, :, :Tb].contiguous(), c[sel].contiguous(), z[sel, :, :Tb].contiguous())
        out[sel, :, :Tb] = o
    return out


def scatter_batch(tensors: Optional[Sequence[torch.Tensor]], counts: Sequence[int], src: int = 0,
                  group=None, device=None, meta: Optional[Sequence[tuple]] = None) -> List[torch.Tensor]:
    """Scatter each tensor's dim-0 slices (sizes ``counts``) from ``src`` with batched point-to-point
    transfers (ranks whose slice is empty take no part).  Non-root ranks pass ``tensors=None`` and
    ``meta`` = [(trailing_shape, dtype), ...]."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    outs, ops = [], []
    if rank == src:
        for t in tensors:
            chunks = torch.split(t, list(counts), dim=0)
            outs.append(chunks[src].contiguous())
            for r in range(world):
                if r != src and counts[r] > 0:
                    ops.append(dist.P2POp(dist.isend, chunks[r].contiguous(), r, group))
    else:
        for trailing, dtype in meta:
            out = torch.empty((counts[rank],) + tuple(trailing), dtype=dtype, device=device)
            outs.append(out)
            if counts[rank] > 0:
                ops.append(dist.P2POp(dist.irecv, out, src, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return outs


def gather_batch(local: torch.Tensor, counts: Sequence[int], dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gather dim-0 slices back on ``dst`` (returns None elsewhere)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ops = []
    if rank == dst:
        parts = []
        for r in range(world):
            if r == dst:
                parts.append(local.contiguous())
            else:
                buf = torch.empty((counts[r],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
                parts.append(buf)
                if counts[r] > 0:
                    ops.append(dist.P2POp(dist.irecv, buf, r, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return torch.cat(parts, dim=0)
    if counts[rank] > 0:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst, group)]):
            w.wait()
    return None


def sharded_solve(solve: Callable[..., torch.Tensor], mu, mask, c, z, *, src: int = 0, group=None, device=None,
                  n_mel: Optional[int] = None, gin: Optional[int] = None, T: Optional[int] = None,
                  batch: Optional[int] = None) -> Optional[torch.Tensor]:
    """Root holds ``mu (B,M,T)``, ``mask (B,1,T)``, ``c (B,gin)``, ``z (B,M,T)``; every rank runs
    ``solve(mu_s, mask_s, c_s, z_s)`` on its contiguous slice; root gets the (B,M,T) result."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if None in (batch, n_mel, T, gin):      # shapes not known on every rank: one tiny broadcast (host-visible)
        if rank == src:
            hdr = torch.tensor([mu.shape[0], mu.shape[1], mu.shape[2], c.shape[1]], dtype=torch.int64, device=mu.device)
        else:
            hdr = torch.empty(4, dtype=torch.int64, device=device)
        dist.broadcast(hdr, src=src, group=group)
        batch, n_mel, T, gin = [int(v) for v in hdr.tolist()]
    counts = split_counts(batch, world)
    f32 = torch.float32
    meta = [((n_mel, T), f32), ((1, T), f32), ((gin,), f32), ((n_mel, T), f32)]
    mu_s, mask_s, c_s, z_s = scatter_batch([mu, mask, c, z] if rank == src else None, counts, src, group, device, meta)
    out = solve(mu_s, mask_s, c_s, z_s) if counts[rank] > 0 else mu_s
    return gather_batch(out, counts, src, group)
