"""Row-sharded scans across the GPUs of one box (one process per GPU).

The path shards by striped rows with no data-path collective:
rank g owns rows [r_g, r_{g+1}) of the GLOBAL striped matrix (+ M-1 halo
rows).  The only exchange is one all-gather of the 272-byte result block per rank
(32 packed per-column winners + the first cell), merged on every rank with
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


class _DeviceWords:
    """__cuda_array_interface__ view of `n` uint64 words at a raw device address."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


class WinnerExchange:
    """The per-scan exchange of a row-sharded argmax, device-resident: the 272-byte result block
    the scan kernels leave in HBM (32 packed per-column winners + the first cell) is all-gathered
    straight from there over NCCL/NVLink on the scan's own stream -- no round trip through host
    memory, no host synchronisation -- and only the gathered blocks (world x 272 bytes) are copied
    to pinned host memory, where `finish` applies the tie rule with ``lm_argmax_merge``.  Several slots,
    so the exchange of scan k overlaps the following scans.

    With the gloo backend (CPU tests) the blocks are host arrays passed to `enqueue` directly."""

    def __init__(self, device=None, group=None, slots=4):
        from . import RESULT_BLOCK_WORDS
        self.group, self.device = group, device
        self.world = dist.get_world_size(group)
        self.words = RESULT_BLOCK_WORDS
        self.cuda = device is not None and torch.device(device).type == "cuda"
        self.all = [torch.empty((self.world, self.words), dtype=torch.int64, device=device) for _ in range(slots)]
        self.host = [torch.empty((self.world, self.words), dtype=torch.int64, pin_memory=self.cuda) for _ in range(slots)]
        self.events = [torch.cuda.Event() for _ in range(slots)] if self.cuda else [None] * slots
        self.side = torch.cuda.Stream(device=device) if self.cuda else None
        self._views = {}

    def enqueue(self, block, slot=0):
        """`block`: device address (int) of this rank's result block, or (CPU) an int64[34] tensor.
        Enqueues all-gather + D2H on the current stream; returns immediately."""
        if isinstance(block, int):
            if block not in self._views:
                self._views[block] = torch.as_tensor(_DeviceWords(block, self.words), device=self.device)
            block = self._views[block]
        if not self.cuda:
            dist.all_gather_into_tensor(self.all[slot], block.reshape(1, self.words), group=self.group)
            self.host[slot].copy_(self.all[slot])
            return
        # the collective is ordered after the scan that produced `block` (NCCL's stream waits for the
        # current stream), but the current stream does NOT wait for it: the next scan starts at once
        # and the gather + copy-out complete beside it on a side stream
        work = dist.all_gather_into_tensor(self.all[slot], block.reshape(1, self.words), group=self.group, async_op=True)
        with torch.cuda.stream(self.side):
            work.wait()
            self.host[slot].copy_(self.all[slot], non_blocking=True)
            self.events[slot].record()

    def finish(self, slot, row_offsets, rows_global, rule=RULE_AVX2):
        """Wait for `slot`'s exchange and merge: the cell one call over the whole matrix returns."""
        from . import decode_keys
        if self.cuda:
            self.events[slot].synchronize()
        blocks = self.host[slot].numpy().view(np.uint64)
        parts = np.stack([decode_keys(blocks[r, :LANES]) for r in range(self.world)])
        first_cell = float(np.array([blocks[0, LANES] & np.uint64(0xFFFFFFFF)], dtype=np.uint64).astype(np.uint32).view(np.float32)[0])
        offsets = np.asarray(list(row_offsets) + [rows_global], dtype=np.uint64)
        return argmax_merge(parts, offsets, first_cell, rows_global, rule)

    def merge(self, local, row_begin, rows_global, rule=RULE_AVX2):
        """Host-input convenience form (one synchronous exchange): `local` = result of
        DeviceSequence.score_argmax(..., with_columns="numpy") on this rank's slab, or None."""
        from . import f32_key
        blk = np.zeros(self.words, dtype=np.uint64)
        if local is not None:
            cols = local["columns"]
            if not isinstance(cols, np.ndarray):
                arr = np.zeros(LANES, dtype=COLWIN_DTYPE)
                for i, (r, sc, v) in enumerate(cols):
                    arr[i] = (r, sc, v)
                cols = arr
            keys = (f32_key(cols["score"]).astype(np.uint64) << np.uint64(32)) | cols["row"].astype(np.uint64)
            blk[:LANES] = np.where(cols["valid"] != 0, keys, np.uint64(0))
            blk[LANES] = np.array([local["first_cell"]], dtype=np.float32).view(np.uint32)[0]
        t = torch.from_numpy(blk.view(np.int64).copy())
        if self.cuda:
            t = t.to(self.device)
        starts = torch.tensor([row_begin], dtype=torch.int64, device=self.device if self.cuda else None)
        all_starts = torch.empty(self.world, dtype=torch.int64, device=starts.device)
        dist.all_gather_into_tensor(all_starts, starts, group=self.group)
        self.enqueue(t, 0)
        return self.finish(0, [int(x) for x in all_starts.cpu().tolist()], rows_global, rule)


def allgather_merge(local, row_begin, rows_global, rule=RULE_AVX2, device=None, group=None):
    """`local` = result of DeviceSequence.score_argmax(..., with_columns=True)
    on this rank's slab (or None for an empty slab).  Returns the global hit."""
    return WinnerExchange(device, group).merge(local, row_begin, rows_global, rule)
