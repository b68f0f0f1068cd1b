"""Row-sharded scans across the GPUs of one box (one process per GPU).

The path shards by striped rows with no data-path collective:
rank g owns rows [r_g, r_{g+1}) of the GLOBAL striped matrix (+ M-1 halo
rows).  The only exchange is one all-gather of 33 small records per rank
(32 per-column winners + the first cell), merged on every rank with
``lm_argmax_merge`` so the result is the cell a single call over the whole
matrix would return (later slabs win row ties, then the back-end's rule).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import LANES, RULE_AVX2, argmax_merge


def shard_rows(rows_global: int, world: int, rank: int):
    """Contiguous row slab of `rank` (balanced to within one row)."""
    return rows_global * rank // world, rows_global * (rank + 1) // world


def pack_winners(columns, first_cell, row_begin, device=None):
    """[33, 3] float64: rows 0..31 = (row, score, valid), row 32 = (row_begin, first_cell, 1).
    float64 holds 32-bit rows and f32 scores exactly."""
    t = torch.empty((LANES + 1, 3), dtype=torch.float64)
    for c, (row, score, valid) in enumerate(columns):
        t[c, 0], t[c, 1], t[c, 2] = float(row), float(score), float(valid)
    t[LANES, 0], t[LANES, 1], t[LANES, 2] = float(row_begin), float(first_cell), 1.0
    return t.to(device) if device is not None else t


def allgather_merge(local, row_begin, rows_global, rule=RULE_AVX2, device=None, group=None):
    """`local` = result of DeviceSequence.score_argmax(..., with_columns=True)
    on this rank's slab (or None for an empty slab).  Returns the global hit."""
    world = dist.get_world_size(group)
    if local is None:
        cols, fc = [(0, 0.0, 0)] * LANES, 0.0
    else:
        cols, fc = local["columns"], local["first_cell"]
    mine = pack_winners(cols, fc, row_begin, device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    out = [o.cpu() for o in out]
    parts = [[(int(x[0]), float(x[1]), int(x[2])) for x in o[:LANES].tolist()] for o in out]
    offsets = [int(o[LANES, 0].item()) for o in out] + [rows_global]
    return argmax_merge(parts, offsets, float(out[0][LANES, 1].item()), rows_global, rule)
