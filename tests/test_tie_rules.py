"""CPU parity of the host-side finishing step of every fused argmax: per-column winners ("column
maximum, LAST row attaining it" -- what the kernels emit) + lm_argmax_merge must give the cell the
oracle's restatement of each reference back-end returns (generic pli/mod.rs:135-155, AVX2
avx2.rs:349-426 incl. its (0,0) start cell, SSE2 sse2.rs:175-261), on matrices full of exact ties,
for one slab and for several row slabs (multi-GPU merge).  Also the packed-key round trip."""
import numpy as np
import pytest

import lightmotif_b200 as lm
from oracle import oracle as O


def winners(scores, r0, r1):
    cols = np.zeros(32, dtype=lm.COLWIN_DTYPE)
    for c in range(32):
        col = scores[r0:r1, c]
        if col.size:
            m = col.max()
            cols[c] = (int(np.flatnonzero(col == m)[-1]), m, 1)
    return cols


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("rule", [0, 1, 2])
def test_merge_reproduces_backend_tie_rules(seed, rule):
    rng = np.random.default_rng(seed)
    rows = int(rng.integers(1, 40))
    scores = rng.integers(-2, 3, size=(rows, 32)).astype(np.float32)  # few distinct values: many ties
    if seed % 3 == 0:
        scores[0, 0] = scores.max()          # AVX2: the start cell (0,0) already holds the maximum
    if seed % 4 == 1:
        scores[:] = -np.inf                  # all -inf
    want = O.argmax(scores, rule)
    # one slab
    got = lm.argmax_merge(winners(scores, 0, rows)[None, :], np.array([0, rows], np.uint64), float(scores[0, 0]), rows, rule)
    assert (got["row"], got["col"]) == want and got["position"] == want[1] * rows + want[0]
    assert np.float32(got["score"]) == scores[want] or (np.isneginf(got["score"]) and np.isneginf(scores[want]))
    # 1..4 row slabs (some possibly empty)
    for world in (2, 3, 4):
        cuts = [rows * r // world for r in range(world + 1)]
        parts = np.stack([winners(scores, cuts[r], cuts[r + 1]) for r in range(world)])
        got = lm.argmax_merge(parts, np.array(cuts, np.uint64), float(scores[0, 0]), rows, rule)
        assert (got["row"], got["col"]) == want, (world, cuts)


def test_packed_keys_round_trip():
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.normal(0, 30, 200).astype(np.float32),
                           np.array([0.0, -0.0, np.inf, -np.inf, 1e-38, -1e-38, 3.4e38, -3.4e38], np.float32)])
    keys = lm.f32_key(vals)
    nz = ~((vals == 0) & np.signbit(vals))  # -0.0 keys just below +0.0; the sequential sums never produce it
    order = np.argsort(vals[nz], kind="stable")
    assert np.all(np.diff(keys[nz][order].astype(np.int64)) >= 0)  # the key preserves the order of the scores
    for chunk in range(0, 192, 32):
        v = vals[chunk:chunk + 32]
        rows_ = rng.integers(0, 2**32 - 1, 32).astype(np.uint64)
        packed = (lm.f32_key(v).astype(np.uint64) << np.uint64(32)) | rows_
        cw = lm.decode_keys(packed)
        assert np.array_equal(cw["score"].view(np.uint32), v.view(np.uint32))
        assert np.array_equal(cw["row"], rows_) and np.all(cw["valid"] == 1)
    assert np.all(lm.decode_keys(np.zeros(32, np.uint64))["valid"] == 0)
