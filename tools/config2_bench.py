#!/usr/bin/env python
"""BASELINE configs[2] through the batch call: 1024 JASPAR motifs x 100 Mbp, per-motif argmax."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
jas = bench.jaspar()[1]
print(json.dumps(bench.configs_block(6555.5, jas)["config2_many_motifs"]))
