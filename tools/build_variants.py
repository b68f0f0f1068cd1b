#!/usr/bin/env python
"""Build experiment variants of the library: tools/_variants/lib_<name>.so, one per
`name=-DFLAG=.. -DFLAG2=..` argument (tools/gpu_sweep.sh runs tools/kbench.py on each)."""
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from lightmotif_b200 import build as B  # noqa: E402


def one(arg):
    name, _, flags = arg.partition("=")
    out = REPO / "tools" / "_variants" / f"lib_{name}.so"
    out.parent.mkdir(exist_ok=True)
    B.build(force=True, extra_flags=flags.split(), out=out)
    return out


if __name__ == "__main__":
    with ThreadPoolExecutor(4) as ex:
        for p in ex.map(one, sys.argv[1:]):
            print(p)
