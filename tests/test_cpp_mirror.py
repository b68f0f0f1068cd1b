"""The C++ host mirror (include/lightmotif_b200.hpp) replays the reference's own
pipeline tests (tests/cpp/test_pipeline.cpp <- lightmotif/tests/{dna,encode,stripe}.rs)
through the C ABI.  CPU: it must compile, link and report UnsupportedBackend
(no CPU fall-back).  GPU: every check must pass."""
import subprocess
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
SRC = REPO / "tests" / "cpp" / "test_pipeline.cpp"
BIN = REPO / "tests" / "cpp" / "test_pipeline"


def build_binary():
    from lightmotif_b200 import build as lm_build
    lib = lm_build.build()
    deps = [SRC, REPO / "include" / "lightmotif_b200.hpp", REPO / "include" / "lightmotif_b200.h", lib]
    if BIN.exists() and all(BIN.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return BIN
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", str(REPO / "include"), str(SRC),
           "-L", str(lib.parent), "-llightmotif_b200", f"-Wl,-rpath,{lib.parent}", "-o", str(BIN)]
    subprocess.run(cmd, check=True)
    return BIN


def test_cpp_mirror_builds_and_has_no_cpu_fallback():
    exe = build_binary()
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    # 77 = Pipeline::cuda() threw UnsupportedBackend (CPU-only box); 0 = ran on a GPU box
    assert p.returncode in (0, 77), p.stdout + p.stderr
    if p.returncode == 77:
        assert "UNSUPPORTED" in p.stdout


def test_cpp_mirror_host_types_match_oracle(kat):
    """CountMatrix -> to_freq -> to_weight -> to_scoring -> to_discrete in the C++ mirror give the
    oracle's (pinned) matrices bit for bit."""
    import numpy as np
    from oracle import oracle as O
    exe = build_binary()
    p = subprocess.run([str(exe), "--host"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = p.stdout.strip().splitlines()
    d = kat["dna"]
    want = O.pssm_from_sequences(O.DNA, d["patterns"], d["pseudocount"])
    m = len(want)
    got = np.array([[int(x, 16) for x in ln.split()] for ln in lines[:m]], dtype=np.uint32)
    assert np.array_equal(got, want[:, :5].view(np.uint32))
    dm = O.Discrete(O.DNA, want)
    got_dm = np.array([[int(x) for x in ln.split()] for ln in lines[m:2 * m]], dtype=np.uint8)
    assert np.array_equal(got_dm, dm.data[:, :5])


@pytest.mark.gpu
def test_cpp_mirror_replays_reference_tests():
    exe = build_binary()
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert " 0 failures" in p.stdout
