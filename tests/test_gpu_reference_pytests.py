"""The reference's own Python tests for the hot path (lightmotif-py/lightmotif/tests/test_pipeline.py,
test_scanner.py, test_sequence.py) run UNMODIFIED against `lightmotif_b200.compat`.

The test files are the reference's and are not part of this repository.  `tests/fetch_reference_pytests.sh`
copies them from /root/reference into tests/_ref_pytests/ (git-ignored scratch, it travels with the
snapshot to the GPU box like built libraries do); without them this test is skipped and
tests/test_gpu_compat.py -- the same assertions, replayed from tests/golden/kat.json -- stands in."""
import importlib.util
import sys
import unittest
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

HERE = Path(__file__).resolve().parent
FILES = ["test_pipeline.py", "test_scanner.py", "test_sequence.py"]


def _candidates():
    for d in (HERE / "_ref_pytests", Path("/root/reference/lightmotif-py/lightmotif/tests")):
        if all((d / f).exists() for f in FILES):
            return d
    return None


@pytest.mark.parametrize("name", FILES)
def test_reference_python_test_file_unmodified(name):
    d = _candidates()
    if d is None:
        pytest.skip("reference test files not present (run tests/fetch_reference_pytests.sh where /root/reference exists)")
    sys.path.insert(0, str(HERE / "reference_alias"))
    try:
        for mod in [m for m in sys.modules if m == "lightmotif" or m.startswith("lightmotif.")]:
            del sys.modules[mod]
        spec = importlib.util.spec_from_file_location("ref_" + name[:-3], d / name)
        module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(module)
        suite = unittest.defaultTestLoader.loadTestsFromModule(module)
        assert suite.countTestCases() > 0
        result = unittest.TextTestRunner(verbosity=0).run(suite)
        assert result.wasSuccessful(), (result.failures, result.errors)
    finally:
        sys.path.remove(str(HERE / "reference_alias"))
        for mod in [m for m in sys.modules if m == "lightmotif" or m.startswith("lightmotif.")]:
            del sys.modules[mod]
