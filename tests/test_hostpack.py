"""Host side of the packed upload (lightmotif_b200/csrc/hostpack.cpp): 4 DNA symbols per byte, little end
first.  Pure host code, so it is checked here without a GPU: the AVX2 path against the scalar path and
against numpy, on aligned and ragged lengths, and the rejection of chunks that hold a symbol > 3 (N)."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def packers():
    from lightmotif_b200 import capi
    lib = capi.lib()
    fns = []
    for sym in ("_ZN3lmi5pack2EPKhmPh", "_ZN3lmi12pack2_scalarEPKhmPh"):
        f = getattr(lib, sym)
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        f.restype = C.c_bool
        fns.append(f)
    return fns


def reference(sym):
    s = sym.reshape(-1, 4).astype(np.uint8)
    return (s[:, 0] | (s[:, 1] << 2) | (s[:, 2] << 4) | (s[:, 3] << 6)).astype(np.uint8)


@pytest.mark.parametrize("n", [4, 8, 12, 124, 128, 132, 252, 256, 4096, 65536 + 20, 1 << 20])
def test_pack2_matches_numpy(packers, n):
    rng = np.random.default_rng(n)
    sym = rng.integers(0, 4, n, dtype=np.uint8)
    want = reference(sym)
    for f in packers:
        out = np.full(n // 4 + 8, 0xEE, dtype=np.uint8)
        assert f(sym.ctypes.data, n, out.ctypes.data)
        assert np.array_equal(out[: n // 4], want)
        assert (out[n // 4:] == 0xEE).all()          # nothing written past the packed bytes


@pytest.mark.parametrize("pos", [0, 3, 31, 127, 128, 1000, 4095])
@pytest.mark.parametrize("bad", [4, 20, 255])
def test_pack2_rejects_symbols_that_do_not_fit(packers, pos, bad):
    sym = np.random.default_rng(pos).integers(0, 4, 4096, dtype=np.uint8)
    sym[pos] = bad
    out = np.zeros(1024, dtype=np.uint8)
    for f in packers:
        assert not f(sym.ctypes.data, sym.size, out.ctypes.data)


def test_pack2_unaligned_source(packers):
    base = np.random.default_rng(5).integers(0, 4, 1 << 16, dtype=np.uint8)
    for off in (1, 3, 17):
        sym = base[off: off + 32768]
        for f in packers:
            out = np.zeros(8192, dtype=np.uint8)
            assert f(sym.ctypes.data, sym.size, out.ctypes.data)
            assert np.array_equal(out, reference(np.ascontiguousarray(sym)))
