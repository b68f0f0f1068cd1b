#!/usr/bin/env python
"""Host -> device upload of a pinned striped DNA matrix through lm_seq_from_striped: plain copy vs the
two-ended upload (plain chunks from the front + 2-bit packed chunks from the back) for several numbers
of packing threads.   python tools/upload_bench.py [--gb 5] [--threads 0,4,8,15]"""
import argparse
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lightmotif_b200 as lm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=5.0)
    ap.add_argument("--threads", default="")
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--thp", action="store_true", help="host matrix in transparent huge pages, pinned with cudaHostRegister")
    a = ap.parse_args()
    cores = len(os.sched_getaffinity(0))
    rows = int(a.gb * 1e9) // 32
    wrap = 19
    if a.thp:
        import mmap
        nbytes = (rows + wrap) * 32
        size = -(-nbytes // (2 << 20)) * (2 << 20)
        mm = mmap.mmap(-1, size, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        mm.madvise(mmap.MADV_HUGEPAGE)
        h = np.frombuffer(mm, dtype=np.uint8, count=nbytes).reshape(rows + wrap, 32)
        h[:] = 0  # fault the pages in (as huge pages where the kernel grants them)
        rc = torch.cuda.cudart().cudaHostRegister(h.ctypes.data, size, 0)
        print("# cudaHostRegister ->", rc, " AnonHugePages:", [l for l in open("/proc/meminfo") if "AnonHugePages" in l][0].strip(),
              " thp:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip())
    else:
        host = torch.empty((rows + wrap, 32), dtype=torch.uint8, pin_memory=True)
        h = host.numpy()
    rng = np.random.default_rng(1)
    blk = rng.integers(0, 4, size=(1 << 20, 32), dtype=np.uint8)
    for r0 in range(0, rows + wrap, 1 << 20):
        n = min(1 << 20, rows + wrap - r0)
        h[r0:r0 + n] = blk[:n]
    pli = lm.Pipeline.cuda(lm.DNA)
    print(f"# {cores} host threads available, matrix {h.nbytes / 1e9:.2f} GB pinned")
    ts = [int(t) for t in (a.threads.split(",") if a.threads else [str(max(1, cores - 1))])]
    combos = [(0, 0)] + [(m, t) for m in (4, 3, 2, 1) for t in ts]
    for mode, thr in combos:
        pli.set_upload(mode, thr)
        ts = []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            s = pli.sequence_from_striped(h, rows * 32, wrap)
            pli.synchronize()
            ts.append(time.perf_counter() - t0)
            up = pli.last_upload
            s.free()
        best = min(ts)
        print(f"mode={mode} threads={thr:3d}  best {best * 1e3:7.1f} ms  median {np.median(ts) * 1e3:7.1f} ms  -> {h.nbytes / best / 1e9:6.1f} Gbp/s"
              f"  pcie {up['pcie_bytes'] / 1e9:.2f} GB  packed/plain chunks {up['packed_chunks']}/{up['plain_chunks']}")


if __name__ == "__main__":
    main()
