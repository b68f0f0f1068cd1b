"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports
every symbol include/lightmotif_b200.h declares (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (REPO / "include" / "lightmotif_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lm_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from lightmotif_b200 import build, capi
    build.build()
    lib = capi.lib()
    names = declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(capi.SYMBOLS) == names
    assert lib.lm_abi_version() == 2
    assert lib.lm_strerror(-3) == b"not enough wrapping rows for motif"


def test_no_gpu_means_unsupported_backend_not_a_fallback():
    """Pipeline::cuda() mirrors Pipeline::avx2() -> Err(UnsupportedBackend)."""
    import torch
    import lightmotif_b200 as lm
    if torch.cuda.is_available():
        return
    try:
        lm.Pipeline.cuda()
    except lm.UnsupportedBackend:
        pass
    else:
        raise AssertionError("expected UnsupportedBackend without a GPU")


def test_host_merge_rules():
    """lm_argmax_merge is pure host code: tie rules over per-column winners."""
    import lightmotif_b200 as lm
    cols = [(0, -1.0, 1)] * 32
    a = list(cols); a[3] = (10, 5.0, 1); a[7] = (2, 5.0, 1)
    b = list(cols); b[3] = (4, 5.0, 1)
    # two slabs of 100 rows; column 3 reaches 5.0 in both, column 7 in slab 0
    g = lm.argmax_merge([a, b], [0, 100, 200], -1.0, 200, lm.RULE_GENERIC)
    assert (g["row"], g["col"]) == (104, 3)          # last row, then last column
    v = lm.argmax_merge([a, b], [0, 100, 200], -1.0, 200, lm.RULE_AVX2)
    assert (v["row"], v["col"]) == (104, 3)          # first column, last row in it
    s = lm.argmax_merge([a, b], [0, 100, 200], -1.0, 200, lm.RULE_SSE2)
    assert (s["row"], s["col"]) == (2, 7)            # last column
    q = lm.argmax_merge([a, b], [0, 100, 200], 5.0, 200, lm.RULE_AVX2)
    assert (q["row"], q["col"]) == (0, 0)            # avx2.rs:411-421 (0,0) quirk
    assert v["position"] == 3 * 200 + 104
