"""Shared fixtures.  GPU tests are marked ``@pytest.mark.gpu``; everything else
runs on the CPU-only container (``pytest -m "not gpu"``)."""
import json
import lzma
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
GOLDEN = REPO / "tests" / "golden"
sys.path.insert(0, str(REPO))


# the reference's own test files (scratch copy, see test_gpu_reference_pytests.py) and the
# `lightmotif` alias package are loaded by that test only, never collected directly
collect_ignore = ["_ref_pytests", "reference_alias"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def kat():
    return json.loads((GOLDEN / "kat.json").read_text())


@pytest.fixture(scope="session")
def ecoli_kat():
    return json.loads((GOLDEN / "ecoli_kat.json").read_text())


def load_ecoli_codes():
    """E. coli K12 genome as symbol codes (A=0 C=1 T=2 G=3), 4 641 652 bp."""
    blob = lzma.decompress((GOLDEN / "ecoli.2bit.xz").read_bytes())
    n = int.from_bytes(blob[:8], "little")
    packed = np.frombuffer(blob, dtype=np.uint8, offset=8)
    codes = np.empty((packed.size, 4), dtype=np.uint8)
    for q in range(4):
        codes[:, q] = (packed >> (2 * q)) & 3
    return codes.reshape(-1)[:n].copy()


@pytest.fixture(scope="session")
def ecoli_codes():
    return load_ecoli_codes()


@pytest.fixture(scope="session")
def ecoli_ascii(ecoli_codes):
    return np.frombuffer(b"ACTG", dtype=np.uint8)[ecoli_codes]


@pytest.fixture(scope="session")
def jaspar():
    z = np.load(GOLDEN / "jaspar2024.npz")
    widths = z["widths"]
    offs = np.concatenate([[0], np.cumsum(widths)])
    motifs = [z["counts"][offs[i]:offs[i + 1]] for i in range(len(widths))]
    return {"rest": z["rest"], "motifs": motifs, "names": z["names"]}


@pytest.fixture(scope="session")
def abyb1():
    return (GOLDEN / "abyB1.txt").read_bytes().strip()
