"""Where does the per-step time of a multi-GPU scan go?  torchrun --nproc-per-node 2 tools/dist_probe.py"""
import os, sys, time
from pathlib import Path
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import lightmotif_b200 as lm
from lightmotif_b200 import dist as lmdist
from bench import rest_pssm, SEED, WIDTH

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per = 5_000_000_000; total = per * world; rows_g = -(-total // 32)
r0, r1 = rows_g * rank // world, rows_g * (rank + 1) // world
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
pli = lm.Pipeline.cuda(lm.DNA, device=local, stream=stream.cuda_stream)
pssm = rest_pssm()
seq = pli.sequence_synthetic(total, WIDTH - 1, seed=SEED, row_begin=r0, n_rows=r1 - r0)
xchg = lmdist.WinnerExchange(device=dev)
offs = [rows_g * r // world for r in range(world)]
K = 20

def timed(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    for _ in range(K): fn()
    e1.record(stream); torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0: print(f"{name:34s} gpu {e0.elapsed_time(e1)/K:.3f} ms/step   host {(t1-t0)*1e3/K:.3f} ms/step", flush=True)

def scan_only():
    seq.score_argmax_launch(pssm, slot=0); seq.score_argmax_finish(slot=0)
def scan_nosync_pair():
    seq.score_argmax_launch(pssm, slot=0); seq.score_argmax_launch(pssm, slot=1)
    seq.score_argmax_finish(slot=0); seq.score_argmax_finish(slot=1)
blk = torch.zeros(34, dtype=torch.int64, device=dev)
def gather_only():
    xchg.enqueue(blk, 0)
def gather_sync():
    xchg.enqueue(blk, 0); xchg.events[0].synchronize()
def scan_gather():
    seq.score_argmax_launch(pssm, slot=0); xchg.enqueue(pli.result_block_ptr(0), 0)
    seq.score_argmax_finish(slot=0); xchg.finish(0, offs, rows_g)
def pipelined(n, gather=True, merge=True):
    def launch(k):
        seq.score_argmax_launch(pssm, slot=k & 1)
        if gather: xchg.enqueue(pli.result_block_ptr(k & 1), slot=k & 1)
    def finish(k):
        seq.score_argmax_finish(slot=k & 1)
        if gather and merge: xchg.finish(k & 1, offs, rows_g)
    launch(0)
    for k in range(n):
        if k + 1 < n: launch(k + 1)
        finish(k)
def timed_block(name, fn):
    fn(3)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    fn(K)
    e1.record(stream); torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0: print(f"{name:34s} gpu {e0.elapsed_time(e1)/K:.3f} ms/step   host {(t1-t0)*1e3/K:.3f} ms/step", flush=True)
timed_block("pipelined scan only", lambda n: pipelined(n, False))
timed_block("pipelined scan + gather (no merge)", lambda n: pipelined(n, True, False))
timed_block("pipelined scan + gather + merge", lambda n: pipelined(n, True, True))
timed("scan only (launch+finish)", scan_only)
timed("2 scans in flight", scan_nosync_pair)
timed("all-gather only (no host sync)", gather_only)
timed("all-gather + event sync", gather_sync)
timed("scan + gather + merge (serial)", scan_gather)
dist.barrier(); dist.destroy_process_group()
