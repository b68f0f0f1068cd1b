"""`lightmotif.lib` (the PyO3 extension module of the reference): same names, CUDA back-end."""
from lightmotif_b200.compat import *  # noqa: F401,F403
from lightmotif_b200.compat import AVX2_SUPPORTED, LIMITED_API  # noqa: F401
