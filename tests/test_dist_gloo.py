"""world_size-2 gloo test (CPU) of the N>1 host logic: slab partitioning,
all-gather of per-column winners, lm_argmax_merge == single-call argmax.
The per-slab winners come from the CPU oracle here (no GPU in this tier)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _slab_winners(want, r0, r1):
    """(column max, LAST row attaining it) of rows [r0, r1) -- what the fused kernel emits."""
    cols = []
    for c in range(32):
        col = want[r0:r1, c]
        if col.size == 0:
            cols.append((0, 0.0, 0))
            continue
        m = col.max()
        cols.append((int(np.flatnonzero(col == m)[-1]), float(m), 1))
    return cols


def _worker(rank, world, port, n, seed, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from lightmotif_b200 import dist as lmdist
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    pssm = rng.integers(-3, 4, size=(9, 5)).astype(np.float32)  # small integers => many exact ties
    pssm[:, 4] = -np.inf
    sym = O.synth_dna(n)
    st = O.stripe(O.DNA, sym, 32, 8)
    want = O.score_f32(pssm, st, n)
    rows = want.shape[0]
    r0, r1 = lmdist.shard_rows(rows, world, rank)
    out = {}
    for rule in (0, 1, 2):
        local = {"columns": _slab_winners(want, r0, r1), "first_cell": float(want[r0, 0])}
        got = lmdist.allgather_merge(local, r0, rows, rule)
        r, c = O.argmax(want, rule)
        out[rule] = (got["row"], got["col"], got["position"], got["score"], r, c, float(want[r, c]), rows)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_argmax_merge_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 20011, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out in results:
        for rule, (row, col, pos, score, r, c, s, rows) in out.items():
            assert (row, col) == (r, c), (rank, rule)
            assert pos == c * rows + r and score == s


def test_shard_rows_partition():
    from lightmotif_b200.dist import shard_rows
    for rows in (0, 1, 7, 156250000, 1250000001):
        for world in (1, 2, 3, 8):
            cuts = [shard_rows(rows, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
