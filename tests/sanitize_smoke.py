"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool memcheck python tests/sanitize_smoke.py"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lightmotif_b200 as lm
from oracle import oracle as O

z = np.load(Path(__file__).resolve().parent.parent / "tests/golden/jaspar2024.npz")
pssm = O.pssm_from_counts(O.DNA, z["rest"], 0.1)
dna = lm.Pipeline.cuda(lm.DNA)
n = 300_007
sym = O.synth_dna(n)
sym[1000:1100] = 4
ascii_ = np.frombuffer(b"ACTGN", np.uint8)[sym]
seq = dna.sequence_from_ascii(ascii_, 19)
st = O.stripe(O.DNA, sym, 32, 19)
assert np.array_equal(seq.download(), st)
want = O.score_f32(pssm, st, n)
assert np.array_equal(seq.score(pssm).view(np.uint32), want.view(np.uint32))
dna.set_boot_rows(512)
for fmode in (1, 2):
    dna.set_prune(2); dna.set_filter_mode(fmode)
    best = seq.score_argmax(pssm)
    r, c = O.argmax(want, O.RULE_AVX2)
    assert (best["row"], best["col"]) == (r, c), (fmode, best, r, c)
    assert dna.last_stats["pruned"]
    hits = seq.score_threshold(pssm, 8.0)
    ro, co = O.threshold(want, 8.0)
    assert [(h["row"], h["col"]) for h in hits] == list(zip(ro.tolist(), co.tolist()))
dna.set_prune(0)
assert seq.score_argmax(pssm)["position"] == best["position"]
dm = O.Discrete(O.DNA, pssm)
u8 = seq.score(dm.data) if False else None
enc = dna.encode(ascii_)
assert np.array_equal(enc, sym)
striped = dna.stripe(sym, 19)
assert np.array_equal(striped, st)
# many-motif job loop (argmax + per-motif thresholds) through the planner
off = np.concatenate([[0], np.cumsum(z["widths"])])
motifs = [z["counts"][off[i]:off[i + 1]] for i in range(len(z["widths"])) if 8 <= z["widths"][i] <= 20][:6]
pssms = [O.pssm_from_counts(O.DNA, c, 0.1) for c in motifs] + [pssm]
dna.set_prune(2)
bb = seq.score_argmax_batch(pssms)
dna.set_prune(0)
assert [b["position"] for b in bb] == [seq.score_argmax(p)["position"] for p in pssms]
dna.set_prune(2)
thr = [b["score"] - 2.0 for b in bb]
hb = seq.score_threshold_batch(pssms, thr, cap=256)
dna.set_prune(0)
assert hb == [seq.score_threshold(p, t) for p, t in zip(pssms, thr)]
# protein: class-coded filter, recheck from the residues in global memory
prot = lm.Pipeline.cuda(lm.PROTEIN)
pp = O.pssm_from_sequences(O.PROTEIN, ["SFKELGFDSLTAVELRNRLAAAT", "AFKELGFDSLAAIQLRNRLLADV", "PSRRLGFDSLTAVELRNQLAAST"], 0.1)[:15]
psym = O.synth_protein(200_003)
psym[::501] = 20
pseq = prot.sequence_from_symbols(psym, 14)
prot.set_boot_rows(512)
prot.set_prune(0)
e = pseq.score_argmax(pp)
eh = pseq.score_threshold(pp, e["score"] - 6.0)
prot.set_prune(2)
assert pseq.score_argmax(pp) == e and prot.last_stats["pruned"]
assert pseq.score_threshold(pp, e["score"] - 6.0) == eh
# dense-candidate instance: scan twice at a low threshold (the second scan has seen the density)
dna.set_prune(2); dna.set_filter_mode(0)
dna.set_prune(0)
low = seq.score_threshold(pssm, 2.0)
dna.set_prune(2)
for _ in range(3):
    assert seq.score_threshold(pssm, 2.0) == low
# both strands through the batch call
dna.set_prune(0)
both_want = []
for p in pssms:
    both_want += [seq.score_argmax(p), seq.score_argmax(lm.reverse_complement(p))]
dna.set_prune(2)
assert seq.score_argmax_batch(pssms, strands=lm.STRAND_BOTH) == both_want
# row-sharded scans through a communicator (three ranks on this device): shared maximum, result
# blocks published by the scan kernels, hit lists gathered through the exchange buffers
comm = lm.Comm.local([0, 0, 0])
for p in comm.pipelines:
    p.set_boot_rows(512)
    p.set_prune(2)
slabs = comm.shard_from_striped(st, n, 19)
r, c = O.argmax(want, O.RULE_AVX2)
for k in range(3):
    g = comm.score_argmax(slabs, pssm)
    assert (g["row"], g["col"]) == (r, c), g
sh = comm.score_threshold(slabs, pssm, 8.0)
ro, co = O.threshold(want, 8.0)
assert [(h["row"], h["col"]) for h in sh] == list(zip(ro.tolist(), co.tolist()))
for sl in slabs:
    sl.free()
comm.close()
# packed upload (host threads pack 2-bit chunks, unpack2_kernel expands them; the chunk holding N goes plain) and
# the exact kernel's prefix-table path (tiles without N)
a = np.random.default_rng(3).integers(0, 4, size=((9 << 20) // 32 + 19, 32), dtype=np.uint8)
a[5, 5] = 4
for mode in (4, 2):
    dna.set_upload(mode, threads=2, min_bytes=0)
    up = dna.sequence_from_striped(a, (a.shape[0] - 19) * 32, 19)
    assert np.array_equal(up.download(), a)
    up.free()
dna.set_upload(1, 0, 64 << 20)
clean = O.synth_dna(100_003)
cseq = dna.sequence_from_symbols(clean, 19)
cst = O.stripe(O.DNA, clean, 32, 19)
assert np.array_equal(cseq.score(pssm).view(np.uint32), O.score_f32(pssm, cst, clean.size).view(np.uint32))
print("sanitize smoke ok", best, e)
