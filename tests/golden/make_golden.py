#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/.

Run HERE (in the container that has /root/reference); the GPU box has no
/root/reference, so everything tests / bench need at run time is written
into small files in this directory:

* ``kat.json``       -- the literal test vectors held by the reference's own
                        tests (sequence, patterns, EXPECTED scores, argmax,
                        threshold hits, encode / stripe literals), each with
                        the file:line it was transcribed from.
* ``ecoli.2bit.xz``  -- lightmotif/benches/ecoli.txt (4 641 652 bp, ACGT only)
                        packed 2 bits / base (A=0 C=1 T=2 G=3, the reference's
                        symbol codes, abc.rs:113-135), xz-compressed.
* ``abyB1.txt``      -- lightmotif/benches/abyB1.txt (5 781 aa).
* ``jaspar2024.npz`` -- count matrices parsed from
                        lightmotif-io/benches/JASPAR2024.pwm: MA0138.3 REST
                        (config 2/5) and the first 1024 records whose width is
                        in 6..=30 (config 3).  Column order A, C, T, G, N.
* ``ecoli_kat.json`` -- derived anchors computed with the oracle on the full
                        genome (NOT pinned by the reference; regression only).
"""
import json
import lzma
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))


def parse_jaspar16(path):
    """Minimal JASPAR-2016 ('>ID name' + 'A [ .. ]' rows) reader."""
    records = []
    name, cols = None, {}
    for line in path.read_text().splitlines():
        if line.startswith(">"):
            if name is not None:
                records.append((name, cols))
            name, cols = line[1:].strip(), {}
        elif line.strip():
            sym = line[0]
            body = line[line.index("[") + 1: line.index("]")]
            cols[sym] = [int(x) for x in body.split()]
    if name is not None:
        records.append((name, cols))
    out = []
    for name, cols in records:
        w = len(cols["A"])
        m = np.zeros((w, 5), dtype=np.uint32)
        for sym, j in (("A", 0), ("C", 1), ("T", 2), ("G", 3)):
            m[:, j] = cols[sym]
        out.append((name, m))
    return out


def main():
    # ---- literal KATs transcribed from the reference's tests
    kat = {
        "dna": {
            "source": "lightmotif/tests/dna.rs:19-38",
            "sequence": "ATGTCCCAACAACGATACCCCGAGCCCATCGCCGTCATCGGCTCGGCATGCAGATTCCCAGGCG",
            "patterns": ["GTTGACCTTATCAAC", "GTTGATCCAGTCAAC"],
            "pseudocount": 0.1,
            "expected": [
                -23.07094, -18.678621, -15.219191, -17.745737,
                -18.678621, -23.07094, -17.745737, -19.611507,
                -27.463257, -29.989803, -14.286304, -26.53037,
                -15.219191, -10.826873, -10.826873, -22.138054,
                -38.774437, -30.922688, -5.50167, -24.003826,
                -18.678621, -15.219191, -35.315006, -17.745737,
                -10.826873, -30.922688, -23.07094, -6.4345555,
                -31.855574, -23.07094, -15.219191, -31.855574,
                -8.961102, -26.53037, -27.463257, -14.286304,
                -15.219191, -26.53037, -23.07094, -18.678621,
                -14.286304, -18.678621, -26.53037, -16.152077,
                -17.745737, -18.678621, -17.745737, -14.286304,
                -30.922688, -18.678621],
            "tolerance": 1e-5,
            "exact_positions": [0, 1],          # tests/dna.rs:57-62
            "argmax": 18,                        # tests/dna.rs:122-139
            "threshold": {"-10.0": [18, 27, 32],  # tests/dna.rs:141-173
                          "-15.0": [10, 13, 14, 18, 24, 27, 32, 35, 40, 47]},
            "readme_score0": -23.07094,          # README.md:81-90
            "scanner_max": [18, -5.50167],       # scan.rs:317-333
        },
        "encode": {
            "source": "lightmotif/tests/encode.rs:10-26",
            "sequence": "ATGTCCCAACAACGATACCCCGAGCCCATCGCCGTCATCGGCTCGGCATGCAGATTCCCAGGCG",
            "expected_letters": "ATGTCCCAACAACGATACCCCGAGCCCATCGCCGTCATCGGCTCGGCATGCAGATTCCCAGGCG",
            "unknowns": "ATGTCCCAACAACGATACCNN..................NNNNNNNNATGCAGATTCCCAGGCG",
            "unknown_char": ".",
            "codes": {"A": 0, "C": 1, "T": 2, "G": 3, "N": 4},  # abc.rs:113-135
        },
        "stripe": {
            "source": "lightmotif/tests/stripe.rs:13-45; lightmotif/src/seq.rs:509-541",
            "s2": "ATGTCCCAACAACGATACCCCGAGCCCATCGCCGTCATCGGCTCGGCATGCAGATTCCCAGGCG",
            "s1_unit": "ATGTCCCAACAACGATACCCCGAGCCCATCGCCGTCATCGGCTCGGCATGCAGATTCCCAGGCG",
            "s1_repeats": 16, "s1_tail": "TTATTAT",   # 16*64+7 = 1031 bp
            "u4": {"sequence": "ATGCA",
                   "matrix": ["AGAN", "TCNN"],
                   "wrap2": ["GANN", "CNNN"]},
        },
        "ecoli_tenth": {
            "source": "lightmotif/tests/argmax.rs:21-53; lightmotif-bench/dna.rs:247,282",
            "n": 464165,
            "patterns": ["GTTGACCTTATCAAC", "GTTGATCCAGTCAAC"],
            "argmax_pipeline": 401649,   # last maximum (max_by), all back-ends
            "argmax_strict_first": 391677,  # naive strict-> loop / bio crate
        },
        "scan": {
            "source": "lightmotif/tests/scan.rs:9-83",
            "patterns": ["GTTGACCTTATCAAC", "GTTGATCCAGTCAAC", "GTTGATCCAGTAAAT"],
            "threshold": 5.0, "wrap": 15,
            "block_sizes": [1, 16, 32, 64, 128, 256],
            "pseudocounts": [0.1, 0.0],
        },
        "protein": {
            "source": "lightmotif/benches/score.rs:165-193 (bench only; parity-UNPINNED)",
            "patterns": ["SFKELGFDSLTAVELRNRLAAAT", "AFKELGFDSLAAIQLRNRLLADV",
                         "PSRRLGFDSLTAVELRNQLAAST", "AFREIGFDSLTAVELRNRLGAAA",
                         "SLMEEGLDSLAAVELGGTLQRDT", "GFFDLGMDSLMAVELRRRIEQGV"],
            "pseudocount": 0.1,
        },
    }
    # make sure the transcription of stripe.rs S1 is right
    s1 = (REF / "lightmotif/tests/stripe.rs").read_text()
    s1 = s1[s1.index('const S1: &str = "') + len('const S1: &str = "'):]
    s1 = s1[: s1.index('"')]
    st = kat["stripe"]
    assert s1 == st["s1_unit"] * st["s1_repeats"] + st["s1_tail"], "S1 transcription"
    (HERE / "kat.json").write_text(json.dumps(kat, indent=1) + "\n")

    # ---- E. coli genome, 2 bits / base
    eco = (REF / "lightmotif/benches/ecoli.txt").read_bytes()
    lut = np.full(256, 255, dtype=np.uint8)
    for ch, code in zip(b"ACTG", range(4)):
        lut[ch] = code
    codes = lut[np.frombuffer(eco, dtype=np.uint8)]
    assert codes.max() < 4
    n = codes.size
    pad = (-n) % 4
    c4 = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    packed = (c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8)
    blob = n.to_bytes(8, "little") + packed.tobytes()
    (HERE / "ecoli.2bit.xz").write_bytes(lzma.compress(blob, preset=9 | lzma.PRESET_EXTREME))

    (HERE / "abyB1.txt").write_bytes((REF / "lightmotif/benches/abyB1.txt").read_bytes())

    # ---- JASPAR 2024 count matrices
    recs = parse_jaspar16(REF / "lightmotif-io/benches/JASPAR2024.pwm")
    assert len(recs) == 2346
    rest = [m for name, m in recs if name.startswith("MA0138.3")][0]
    assert rest.shape == (20, 5)
    sel = [(name, m) for name, m in recs if 6 <= m.shape[0] <= 30][:1024]
    assert len(sel) == 1024
    widths = np.array([m.shape[0] for _, m in sel], dtype=np.int32)
    flat = np.concatenate([m for _, m in sel], axis=0)
    np.savez_compressed(HERE / "jaspar2024.npz", rest=rest, widths=widths, counts=flat,
                        names=np.array([n for n, _ in sel]))

    # ---- derived anchors on the full genome (oracle; not reference-pinned)
    from oracle import oracle as O
    pssm = O.pssm_from_sequences(O.DNA, kat["dna"]["patterns"], 0.1)
    st = O.stripe(O.DNA, codes, 32, wrap=14)
    sc = O.score_f32(pssm, st, n)
    u = O.unstripe(sc, n + 1 - 15)
    derived = {
        "note": "computed by oracle/lm_oracle.c, NOT pinned by the reference",
        "full_genome_max": float(u.max()),
        "full_genome_argmax_last": int(np.flatnonzero(u == u.max())[-1]),
        "count_ge_10": int((u >= 10.0).sum()),
    }
    p3 = O.pssm_from_sequences(O.DNA, kat["scan"]["patterns"], 0.1)
    st3 = O.stripe(O.DNA, codes, 32, wrap=15)
    u3 = O.unstripe(O.score_f32(p3, st3, n, wrap=15), n + 1 - 15)
    derived["scan_hits_pc0.1_thr5"] = int((u3 >= 5.0).sum())
    (HERE / "ecoli_kat.json").write_text(json.dumps(derived, indent=1) + "\n")
    print(derived)


if __name__ == "__main__":
    main()
