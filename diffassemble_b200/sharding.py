"""Batch sharding of independent piece-graphs across the GPUs of one node.

Sampling is embarrassingly parallel over graphs (attention never crosses ``batch``
boundaries because PyG batching offsets ``edge_index`` per graph), so each rank takes a
contiguous block of graphs, runs the whole sampling loop on it with NO data-path
collective, and the predicted poses are exchanged once at the end with a single
all-gather.  The reference gets the same effect from Lightning DDP
(``train_script.py:214-219``), which never communicates during sampling either.

Note (reference bug): with ``architecture="exophormer"`` and
``virt_nodes > 0`` the reference's virtual-edge wiring couples the graphs of one batch, so
a rank's result equals the reference run on that rank's sub-batch.
"""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_graphs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [g0, g1) of graphs owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(num_graphs, world_size)
    g0 = rank * base + min(rank, rem)
    return g0, g0 + base + (1 if rank < rem else 0)


def shard_batch(edge_index: torch.Tensor, batch: torch.Tensor, node_tensors: Sequence[torch.Tensor], world_size: int,
                rank: int):
    """Slice a PyG-style batch down to the graphs of ``rank``.

    Returns ``(edge_index_local, batch_local, node_tensors_local, (n0, n1))`` with node ids
    and graph ids re-based to zero.  ``batch`` must be non-decreasing (PyG collation order).
    """
    num_graphs = int(batch.max()) + 1
    g0, g1 = shard_bounds(num_graphs, world_size, rank)
    counts = torch.bincount(batch, minlength=num_graphs)
    starts = torch.cumsum(counts, 0) - counts
    n0 = int(starts[g0]) if g0 < num_graphs else int(counts.sum())
    n1 = int(starts[g1 - 1] + counts[g1 - 1]) if g1 > g0 else n0
    dst = edge_index[1]
    keep = (dst >= n0) & (dst < n1)
    src_kept = edge_index[0][keep]
    if bool(((src_kept < n0) | (src_kept >= n1)).any()):
        raise ValueError("edge_index couples graphs of different shards; the batch cannot be split by graph")
    ei = edge_index[:, keep] - n0
    return ei.contiguous(), (batch[n0:n1] - g0).contiguous(), [t[n0:n1].contiguous() for t in node_tensors], (n0, n1)


def gather_poses(local: torch.Tensor, counts: Sequence[int], group=None) -> torch.Tensor:
    """The one collective of the sampling path: all-gather the final ``[nodes_r, C]`` poses.

    ``counts[r]`` = number of nodes on rank r.  Ragged shards are padded to the largest."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    mx = max(counts)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0],) + tuple(local.shape[1:]))])
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad.contiguous(), group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)])
