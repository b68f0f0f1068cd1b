"""`import lightmotif` -> lightmotif_b200.compat, so that the reference's own Python tests
(lightmotif-py/lightmotif/tests/*.py) run UNMODIFIED against the CUDA back-end: see
tests/test_gpu_reference_pytests.py.  Only on sys.path while that test runs."""
from lightmotif_b200.compat import *  # noqa: F401,F403
from lightmotif_b200.compat import (CountMatrix, EncodedSequence, Hit, Motif, Scanner, ScoringMatrix,  # noqa: F401
                                    StripedScores, StripedSequence, WeightMatrix, create, load, scan, stripe)
from . import lib  # noqa: F401
