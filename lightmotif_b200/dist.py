"""Row-sharded scans across the GPUs of one box (one process per GPU).

The path shards by striped rows with no data-path collective:
rank g owns rows [r_g, r_{g+1}) of the GLOBAL striped matrix (+ M-1 halo
rows).  The only exchange is one all-gather of 33 small records per rank
(32 per-column winners + the first cell), merged on every rank with
``lm_argmax_merge`` so the result is the cell a single call over the whole
matrix would return (later slabs win row ties, then the back-end's rule).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import COLWIN_DTYPE, LANES, RULE_AVX2, argmax_merge


def shard_rows(rows_global: int, world: int, rank: int):
    """Contiguous row slab of `rank` (balanced to within one row)."""
    return rows_global * rank // world, rows_global * (rank + 1) // world


def pack_winners(columns, first_cell, row_begin):
    """[33, 3] float64 (numpy): rows 0..31 = (row, score, valid), row 32 =
    (row_begin, first_cell, 1).  float64 holds 32-bit rows and f32 scores exactly."""
    t = np.zeros((LANES + 1, 3), dtype=np.float64)
    if isinstance(columns, np.ndarray):
        t[:LANES, 0], t[:LANES, 1], t[:LANES, 2] = columns["row"], columns["score"], columns["valid"]
    else:
        t[:LANES] = np.asarray(columns, dtype=np.float64)
    t[LANES] = (float(row_begin), float(first_cell), 1.0)
    return t


class WinnerExchange:
    """Preallocated buffers for the per-step all-gather of the winners
    (33 x 3 doubles per rank; latency-bound, so no per-step allocations)."""

    def __init__(self, device=None, group=None):
        self.group, self.device = group, device
        self.world = dist.get_world_size(group)
        self.mine = torch.empty((LANES + 1, 3), dtype=torch.float64, device=device)
        self.all = torch.empty((self.world * (LANES + 1), 3), dtype=torch.float64, device=device)
        pin = device is not None and torch.device(device).type == "cuda"
        self.host_in = torch.empty((LANES + 1, 3), dtype=torch.float64, pin_memory=pin)
        self.host_out = torch.empty((self.world * (LANES + 1), 3), dtype=torch.float64, pin_memory=pin)

    def merge(self, local, row_begin, rows_global, rule=RULE_AVX2):
        if local is None:
            packed = pack_winners(np.zeros((LANES, 3)), 0.0, row_begin)
        else:
            packed = pack_winners(local["columns"], local["first_cell"], row_begin)
        self.host_in.numpy()[:] = packed
        self.mine.copy_(self.host_in, non_blocking=True)
        dist.all_gather_into_tensor(self.all, self.mine, group=self.group)
        self.host_out.copy_(self.all, non_blocking=False)
        out = self.host_out.numpy().reshape(self.world, LANES + 1, 3)
        parts = np.zeros((self.world, LANES), dtype=COLWIN_DTYPE)
        parts["row"] = out[:, :LANES, 0].astype(np.uint64)
        parts["score"] = out[:, :LANES, 1].astype(np.float32)
        parts["valid"] = out[:, :LANES, 2].astype(np.uint32)
        offsets = np.concatenate([out[:, LANES, 0], [rows_global]]).astype(np.uint64)
        return argmax_merge(parts, offsets, float(out[0, LANES, 1]), rows_global, rule)


def allgather_merge(local, row_begin, rows_global, rule=RULE_AVX2, device=None, group=None):
    """`local` = result of DeviceSequence.score_argmax(..., with_columns=True)
    on this rank's slab (or None for an empty slab).  Returns the global hit."""
    return WinnerExchange(device, group).merge(local, row_begin, rows_global, rule)
