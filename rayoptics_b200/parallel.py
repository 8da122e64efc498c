"""Multi-GPU sharding of pupil grids: one process per GPU, no data-path collective.

Every ray is independent and the surface table is < 32 KB, so each rank
replicates the table, takes a contiguous slice of the grid's chunk space
(field-major, then wavelength, then pupil rows -- SURVEY.md 8(e)), generates its
own start rays on the device and traces them.  The single collective is an
all-gather of the ``[n_tiles, 16]`` partial spot sums (NCCL over NVLink on GPUs;
gloo in the CPU tests), followed by a local combine on every rank.

The reference has no multi-process code at all; there is nothing to mirror.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .engine import combine_summaries


def shard_chunks(n_chunks, rank, world_size):
    """Balanced contiguous split of ``[0, n_chunks)``: the first ``n_chunks %
    world_size`` ranks get one extra chunk.  Returns ``(begin, end)``."""
    base, extra = divmod(int(n_chunks), int(world_size))
    begin = rank*base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_chunks_weighted(chunks_per_tile, tile_weights, rank, world_size):
    """Contiguous split of the chunk space into ``world_size`` ranges of (nearly) equal WORK
    instead of equal length.  ``tile_weights[t]``: relative cost of one chunk of tile ``t``
    (e.g. from the spot sums of an earlier pass: rays that reach the image cost a full trace,
    clipped rays a fraction -- outer fields of an unvignetted specification are cheaper than the
    axial one, and an equal-length split leaves the ranks that hold them idle).  Every rank
    computes the same cut points from the same weights.  Returns ``(begin, end)``."""
    import numpy as np
    w = np.repeat(np.maximum(np.asarray(tile_weights, dtype=np.float64), 1e-12), int(chunks_per_tile))
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [int(np.searchsorted(cum, cum[-1]*r/world_size, side='left')) for r in range(world_size + 1)]
    cuts[0], cuts[-1] = 0, len(w)
    for r in range(1, world_size + 1):          # monotone, never empty where avoidable
        cuts[r] = max(cuts[r], cuts[r - 1])
    return cuts[rank], cuts[rank + 1]


def weights_from_summary(summary, clipped_cost=0.3):
    """Per-tile chunk weights from a combined ``[n_tiles, 16]`` summary: rays that arrive count 1,
    rays that fail ``clipped_cost`` (they stop part way)."""
    s = summary.detach().cpu().numpy() if torch.is_tensor(summary) else summary
    ok = s[:, 0]
    fail = s[:, 1:5].sum(axis=1)
    return ok + clipped_cost*fail


class PendingSummary:
    """Handle of an in-flight all-gather (``gather_summaries(..., async_op=True)``): the
    collective runs on NCCL's own stream while the next grid is traced; ``result()`` makes the
    current stream wait for it and combines the ranks' rows with ONE ``rt_combine_summaries``
    launch (CPU / gloo: torch)."""

    def __init__(self, work, out, shape, world, keep=None):
        self.work, self.out, self.shape, self.world = work, out, shape, world
        self._keep = keep      # the input tensor stays referenced until the collective is done

    def result(self):
        self.work.wait()
        return combine_summaries(self.out.view((self.world,) + self.shape))


def gather_summaries(partial, group=None, async_op=False):
    """All-gather the per-rank partial summaries and combine them.

    ``partial``: ``[n_tiles, 16]`` float64 tensor (CUDA for NCCL, CPU for gloo).
    Returns the combined ``[n_tiles, 16]`` summary, identical on every rank
    (or a ``PendingSummary`` when ``async_op``).  The combine is one
    ``rt_combine_summaries`` launch on CUDA tensors."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if async_op:
            class _Done:
                def result(self_inner):
                    return partial.clone()
            return _Done()
        return partial.clone()
    world = dist.get_world_size(group)
    partial = partial.contiguous()
    shape = tuple(partial.shape)
    # concatenated layout [world*n_tiles, 16]: accepted by both NCCL and gloo
    out = torch.empty((world*shape[0],) + shape[1:], dtype=partial.dtype, device=partial.device)
    if async_op:
        work = dist.all_gather_into_tensor(out, partial, group=group, async_op=True)
        return PendingSummary(work, out, shape, world, keep=partial)
    dist.all_gather_into_tensor(out, partial, group=group)
    return combine_summaries(out.view((world,) + shape))


def trace_grid_sharded(table, grid, group=None, **kwargs):
    """Trace this rank's shard of ``grid`` and return ``(local result,
    combined summary)``.  ``kwargs`` go to ``engine.trace_grid``."""
    from .engine import trace_grid
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    begin, end = shard_chunks(grid.n_chunks, rank, world)
    res = trace_grid(table, grid, begin, end, **kwargs)
    combined = gather_summaries(res.summary, group) if res.summary is not None else None
    return res, combined
