"""GPU parity tests of the row-sharded (multi-GPU) calls of the C ABI: lm_comm_*.

A communicator may list the same device several times, so the whole multi-rank machinery -- shared
running maximum through peer-mapped words, publication of the result blocks by the scan kernel,
arrival flags, gather of the hit lists -- is exercised on a one-GPU box too; with two or more GPUs
the same tests also run across real peers.  The one-process-per-GPU flavour (cudaIpc handles
exchanged through torch.distributed) is covered by a 2-process test.

The checker is the CPU oracle at sizes it finishes in seconds, and the single-call result of the
library itself (already pinned against the oracle in test_gpu_parity / test_gpu_prune) above that."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

pytestmark = pytest.mark.gpu


def _device_sets():
    import torch
    n = torch.cuda.device_count()
    sets = [[0, 0], [0, 0, 0]]
    if n >= 2:
        sets += [[0, 1], list(range(min(n, 4)))]
    return sets


def _tie_pssm(seed, m=9):
    rng = np.random.default_rng(seed)
    pssm = rng.integers(-3, 4, size=(m, 5)).astype(np.float32)  # small integers => many exact ties
    pssm[:, 4] = -np.inf
    return pssm


def _rest(jaspar):
    from lightmotif_b200 import pwm
    return pwm.pssm_from_counts(pwm.DNA, jaspar["rest"], 0.1)


@pytest.mark.parametrize("devices", _device_sets())
@pytest.mark.parametrize("prune", [0, 2])
def test_sharded_argmax_equals_oracle_and_single_call(devices, prune, jaspar):
    import lightmotif_b200 as lm
    from oracle import oracle as O
    n = 3_000_000
    for pssm in (_tie_pssm(1), _rest(jaspar)):
        wrap = len(pssm) - 1
        want = O.score_f32(pssm, O.stripe(O.DNA, O.synth_dna(n), 32, wrap), n)
        comm = lm.Comm.local(devices)
        for p in comm.pipelines:
            p.set_prune(prune)
            p.set_boot_rows(2048)
        slabs = comm.sequence_synthetic(n, wrap)
        assert sum(s.rows for s in slabs) == want.shape[0]
        for rule in (lm.RULE_GENERIC, lm.RULE_AVX2, lm.RULE_SSE2):
            got = comm.score_argmax(slabs, pssm, rule)
            r, c = O.argmax(want, rule)
            assert (got["row"], got["col"]) == (r, c), (rule, got, r, c)
            assert np.float32(got["score"]) == want[r, c]
            assert got["position"] == c * want.shape[0] + r
        if prune:
            assert all(p.last_stats["pruned"] for p in comm.pipelines)
        for s in slabs:
            s.free()
        comm.close()


@pytest.mark.parametrize("devices", _device_sets()[:1] + _device_sets()[2:3])
def test_sharded_argmax_pipelined_slots(devices, jaspar):
    """4 sharded scans in flight (different motifs per slot), as bench.py drives them."""
    import lightmotif_b200 as lm
    from lightmotif_b200 import pwm
    n = 40_000_000
    motifs = [pwm.pssm_from_counts(pwm.DNA, c, 0.1) for c in jaspar["motifs"] if 10 <= len(c) <= 20][:8]
    wrap = max(len(p) for p in motifs) - 1
    single = lm.Pipeline.cuda(lm.DNA)
    whole = single.sequence_synthetic(n, wrap)
    want = [whole.score_argmax(p, rule=lm.RULE_AVX2) for p in motifs]
    comm = lm.Comm.local(devices)
    for p in comm.pipelines:
        p.set_boot_rows(8192)
        p.set_prune(2)
    slabs = comm.sequence_synthetic(n, wrap)
    got = [None] * len(motifs)
    for k in range(len(motifs) + 3):
        if k < len(motifs):
            comm.score_argmax_launch(slabs, motifs[k], slot=k % 4)
        if k >= 3:
            got[k - 3] = comm.score_argmax_finish(slot=(k - 3) % 4, rule=lm.RULE_AVX2)
    assert got == want
    for s in slabs:
        s.free()
    whole.free()
    comm.close()


@pytest.mark.parametrize("devices", _device_sets())
def test_sharded_threshold_equals_oracle(devices, jaspar):
    import lightmotif_b200 as lm
    from oracle import oracle as O
    n = 2_500_000
    pssm = _rest(jaspar)
    wrap = len(pssm) - 1
    want = O.score_f32(pssm, O.stripe(O.DNA, O.synth_dna(n), 32, wrap), n)
    comm = lm.Comm.local(devices)
    slabs = comm.sequence_synthetic(n, wrap)
    for prune in (0, 2):
        for p in comm.pipelines:
            p.set_prune(prune)
        for t in (8.0, 2.0):
            hits, per_rank = comm.score_threshold(slabs, pssm, t, with_counts=True)
            ro, co = O.threshold(want, t)
            assert [(h["row"], h["col"]) for h in hits] == list(zip(ro.tolist(), co.tolist()))
            assert all(np.float32(h["score"]) == want[h["row"], h["col"]] for h in hits)
            assert all(h["position"] == h["col"] * want.shape[0] + h["row"] for h in hits)
            assert sum(per_rank) == len(hits) and len(per_rank) == len(devices)
    for s in slabs:
        s.free()
    comm.close()


@pytest.mark.parametrize("devices", _device_sets()[:2])
def test_shard_from_host_striped_and_tiny_inputs(devices):
    """Slabs cut from a host StripedSequence (halo = the next rows of the matrix), including
    sequences so short that some ranks own no rows at all."""
    import lightmotif_b200 as lm
    from oracle import oracle as O
    pssm = _tie_pssm(3, m=5)
    comm = lm.Comm.local(devices)
    for n in (37, 64, 1000, 77_777):
        rng = np.random.default_rng(n)
        sym = rng.integers(0, 5, size=n).astype(np.uint8)
        st = O.stripe(O.DNA, sym, 32, len(pssm) - 1)
        want = O.score_f32(pssm, st, n)
        slabs = comm.shard_from_striped(st, n, len(pssm) - 1)
        for rule in (lm.RULE_GENERIC, lm.RULE_AVX2, lm.RULE_SSE2):
            got = comm.score_argmax(slabs, pssm, rule)
            r, c = O.argmax(want, rule)
            assert (got["row"], got["col"], np.float32(got["score"])) == (r, c, want[r, c]), (n, rule)
        hits = comm.score_threshold(slabs, pssm, 3.0)
        ro, co = O.threshold(want, 3.0)
        assert [(h["row"], h["col"]) for h in hits] == list(zip(ro.tolist(), co.tolist()))
        for s in slabs:
            s.free()
    # a motif longer than the sequence: None, as the single call
    slabs = comm.shard_from_striped(O.stripe(O.DNA, np.zeros(3, np.uint8), 32, 4), 3, 4)
    assert comm.score_argmax(slabs, pssm) is None
    assert comm.score_threshold(slabs, pssm, 0.0) == []
    comm.close()


def test_slab_handles_refuse_to_grow_wrap():
    """growing wrap on a row slab would overwrite its halo with the local closed form."""
    import lightmotif_b200 as lm
    pli = lm.Pipeline.cuda(lm.DNA)
    slab = pli.sequence_synthetic(1 << 20, 7, row_begin=100, n_rows=1000)
    with pytest.raises(lm.LightmotifError):
        slab.configure_wrap(19)
    whole = pli.sequence_synthetic(1 << 20, 7)
    whole.configure_wrap(19)  # a whole sequence still grows
    assert whole.wrap == 19


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ipc_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import lightmotif_b200 as lm
    from lightmotif_b200 import pwm
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(REPO / "tests" / "golden" / "jaspar2024.npz")
    pssm = pwm.pssm_from_counts(pwm.DNA, z["rest"], 0.1)
    pli = lm.Pipeline.cuda(lm.DNA, device=dev)
    pli.set_boot_rows(8192)
    pli.set_prune(2)
    comm = lm.Comm.from_process_group(pli)
    slabs = comm.sequence_synthetic(n, len(pssm) - 1)
    out = {"argmax": [], "hits": None}
    for k in range(6):  # several scans: slot reuse, epochs, both parities
        comm.score_argmax_launch(slabs, pssm, slot=k % 4)
        out["argmax"].append(comm.score_argmax_finish(slot=k % 4, rule=lm.RULE_AVX2))
    for t in (9.0, 6.0):
        out["hits"] = comm.score_threshold(slabs, pssm, t)
    if rank == 0:
        whole = pli.sequence_synthetic(n, len(pssm) - 1)
        out["want_argmax"] = whole.score_argmax(pssm, rule=lm.RULE_AVX2)
        out["want_hits"] = whole.score_threshold(pssm, 6.0)
    q.put((rank, out))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


def test_one_process_per_gpu_cuda_ipc():
    """lm_comm_init_rank / lm_comm_connect: two processes (on two GPUs when the box has them,
    else sharing GPU 0), exchange buffers shared through cudaIpc handles."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, world, n = _free_port(), 2, 30_000_000
    procs = [ctx.Process(target=_ipc_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = res[0]["want_argmax"]
    for r in range(world):
        assert all(a == want for a in res[r]["argmax"]), (r, res[r]["argmax"], want)
        assert res[r]["hits"] == res[0]["want_hits"]
    assert len(res[0]["want_hits"]) > 0
