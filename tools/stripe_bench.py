#!/usr/bin/env python
"""Encode + stripe + wrap from ASCII (lm_seq_from_ascii, stripe_kernel<true>) on 1 Gbp: kernel time."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lightmotif_b200 as lm  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
rng = np.random.default_rng(1)
ascii_ = np.frombuffer(b"ACTG", np.uint8)[rng.integers(0, 4, min(n, 64_000_000))]
ascii_ = np.tile(ascii_, n // ascii_.size + 1)[:n].copy()
pli = lm.Pipeline.cuda(lm.DNA)
ts = []
for _ in range(4):
    s = pli.sequence_from_ascii(ascii_, 19)
    ts.append(pli.last_kernel_ms)
    s.free()
ms = min(ts)
print(f"stripe_kernel<ascii>: {n / 1e9:g} Gbp  best {ms:.3f} ms -> {n / ms / 1e6:.1f} Gbp/s = {2 * n / ms / 1e6:.1f} GB/s (1 B read + 1 B written)")
