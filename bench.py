#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 PWM scanning engine.

Metric (BASELINE.json): Gbp/s scanned with a width-20 DNA PSSM (JASPAR
MA0138.3 REST), fused score + argmax, on 5 Gbp of synthetic DNA per GPU
(configs[1]; N GPUs scan N x 5 Gbp row-sharded = configs[4] at N = 8, weak
scaling), reported with the HBM roofline of the scoring kernel and the CPU
baseline (AVX2 restatement of the reference's path) timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun ... bench.py --gpus N ...            (one rank per GPU, NCCL)

One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

WIDTH = 20
PER_GPU_BP = 5_000_000_000
SEED = 0x5EED
METRIC = "Gbp/s scanned (width-20 DNA PSSM, fused score+argmax)"


def rest_pssm():
    """MA0138.3 REST counts -> to_freq(0.1).to_scoring(None): host-side O(M*K) set-up with the
    harness's own constructors (lightmotif_b200/pwm.py, same libm log2f as the reference); the count
    matrix comes from the JASPAR fixture."""
    from lightmotif_b200 import pwm
    z = np.load(REPO / "tests" / "golden" / "jaspar2024.npz")
    return pwm.pssm_from_counts(pwm.DNA, z["rest"], 0.1)


def measured_traffic(per_gpu_bp, pruned):
    """DRAM bytes of one launch of the dominant kernel from the committed ncu --set full capture
    (profiles/ncu_traffic.json); only when the capture is of this workload."""
    p = REPO / "profiles" / "ncu_traffic.json"
    if not (p.exists() and pruned):
        return None
    d = json.loads(p.read_text())
    if int(d.get("launch_bp", 0)) != int(per_gpu_bp):
        return None
    return int(d["dram_bytes_read"]) + int(d["dram_bytes_write"])


def measured_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons for the timed region.

    * `sm_mhz` is measured ON THE DEVICE, every step: the scan kernel reads the SM cycle counter and
      the global nanosecond timer at its start and end (lm_ctx_last_sm_clock_mhz), so it is the
      clock the timed kernels actually ran at.
    * throttle reasons (and an NVML clock reading for cross-checking) are sampled through NVML
      immediately before the first and after the last timed step -- the GPU is hot from the warm-up
      steps -- and, on a single GPU, also inline between steps.  NVML calls are kept out of the
      multi-GPU timed region: each one stalls the NCCL-coupled pipeline of both ranks by ~4 ms on
      these hosts (measured: 10 steps went from 1.25 to 2.57 ms/step with three inline queries; a
      2 ms polling thread cost 40 %)."""

    BITS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}

    def __init__(self, index):
        self.index, self.kernel_mhz, self.nvml_mhz, self.reasons, self.max_mhz, self._nvml = index, [], [], set(), None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._nvml = (pynvml, h)
        except Exception:
            self._nvml = None

    def sample_nvml(self):
        if self._nvml is None:
            return
        nv, h = self._nvml
        try:
            self.nvml_mhz.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for name, bit in self.BITS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def sample_kernel(self, mhz):
        if mhz and mhz > 0:
            self.kernel_mhz.append(float(mhz))

    def result(self):
        if self.max_mhz is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=10).stdout
                a, b = [float(x) for x in out.strip().split(",")]
                self.nvml_mhz, self.max_mhz = [a], b
            except Exception:
                pass
        src = self.kernel_mhz or self.nvml_mhz
        return {"sm_mhz": statistics.median(src) if src else None, "sm_max_mhz": self.max_mhz,
                "samples": len(src), "reasons": sorted(self.reasons),
                "how": ("in-kernel cycle counter / global timer, one sample per timed step" if self.kernel_mhz
                        else "NVML") + "; reasons via NVML just before / after the timed steps",
                "nvml_sm_mhz": self.nvml_mhz}


def cpu_baseline(pssm, threads=None, target_s=10.0, sample_bp=256_000_000):
    """The reference's AVX2 path (restated in oracle/lm_avx2_ref.c, the Rust
    crate cannot be built here) on the host cores: materialise StripedScores,
    then argmax -- pssm.score(&striped) + scores.argmax() -- row range split
    over `threads` workers, repeated on a bounded sample for ~target_s."""
    from oracle import oracle as O
    threads = threads or os.cpu_count() or 1
    sym = O.synth_dna(sample_bp, 0, SEED)
    striped = O.stripe(O.DNA, sym, 32, WIDTH - 1)
    rows = striped.shape[0] - (WIDTH - 1)
    scratch = O.aligned_empty((rows, 32), np.float32)
    scratch[:] = 0  # fault the pages in before timing
    O.avx2_score_argmax_mt(pssm, striped, rows, threads, scratch)  # warm-up pass
    passes, t0 = 0, time.perf_counter()
    while True:
        res = O.avx2_score_argmax_mt(pssm, striped, rows, threads, scratch)
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= target_s or passes >= 4096:
            break
    gbps = passes * sample_bp / dt / 1e9
    return {"value": gbps, "unit": "Gbp/s", "cores": threads, "kind": "port",
            "sample": f"{passes} passes over the first {sample_bp / 1e6:.0f} Mbp of the workload "
                      f"({dt:.1f} s wall), AVX2 score (materialised) + argmax, {threads} threads",
            "argmax": [int(res[1]), int(res[2])]}, dt / passes


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pssm = rest_pssm()
    # each "step" = one pass over a bounded sample; K steps + W warm-up
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    sample_bp = 256_000_000
    sym = O.synth_dna(sample_bp, 0, SEED)
    striped = O.stripe(O.DNA, sym, 32, WIDTH - 1)
    rows = striped.shape[0] - (WIDTH - 1)
    scratch = O.aligned_empty((rows, 32), np.float32)
    scratch[:] = 0
    for _ in range(args.warmup):
        O.avx2_score_argmax_mt(pssm, striped, rows, threads, scratch)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.avx2_score_argmax_mt(pssm, striped, rows, threads, scratch)
    dt = time.perf_counter() - t0
    gbps = args.steps * sample_bp / dt / 1e9
    sample = (f"each step = the first {sample_bp / 1e6:.0f} Mbp of the 5 Gbp workload; AVX2 restatement of "
              f"lightmotif's score+argmax (oracle/lm_avx2_ref.c; the Rust crate cannot be built: no rustc), "
              f"{threads} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": gbps, "unit": "Gbp/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "width-20 DNA PSSM (MA0138.3 REST) over 5 Gbp synthetic DNA, score+argmax",
                   "sample_bp_per_step": sample_bp},
        "cpu_baseline": {"value": gbps, "unit": "Gbp/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gbps, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import lightmotif_b200 as lm
    from lightmotif_b200 import dist as lmdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set: keep stdout to the one JSON line
        if "LM_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["LM_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        dist.init_process_group("nccl", device_id=dev)

    per_gpu = args.bp_per_gpu
    total = per_gpu * world
    rows_global = -(-total // 32)
    r0 = rows_global * rank // world
    r1 = rows_global * (rank + 1) // world
    pssm = rest_pssm()
    stream = torch.cuda.Stream(device=dev)  # the library enqueues on this stream; events are recorded on it
    torch.cuda.set_stream(stream)
    pli = lm.Pipeline.cuda(lm.DNA, device=local, stream=stream.cuda_stream)
    # this rank's row slab of the global striped matrix, generated in HBM
    seq = pli.sequence_synthetic(total, WIDTH - 1, seed=SEED, row_begin=r0, n_rows=r1 - r0)
    positions = (r1 - r0) * 32  # cells scored per step on this rank (1 byte each)

    xchg = lmdist.WinnerExchange(device=dev) if world > 1 else None
    row_offsets = [rows_global * r // world for r in range(world)]
    if world > 1 and args.reserve_sms:
        pli.set_reserve_sms(args.reserve_sms)  # room for the NCCL kernel of scan k beside scan k + 1

    # One step = one fused scan of this rank's slab + (N > 1) the exchange of the winners: an NCCL
    # all-gather of the 272-byte result block straight from HBM, enqueued on the scan's stream, and
    # the tie-rule merge on the host.  Steps are software-pipelined `depth` deep: scans k+1 .. k+depth
    # are enqueued before the host waits for scan k (so a late rank or a slow host thread does not
    # idle the other GPUs).
    NSLOT = 4
    depth = args.depth  # scans enqueued ahead of the one the host waits for (< NSLOT)

    def launch(k):
        seq.score_argmax_launch(pssm, slot=k % NSLOT)
        if world > 1:
            xchg.enqueue(pli.result_block_ptr(k % NSLOT), slot=k % NSLOT)

    def finish(k):
        res = seq.score_argmax_finish(slot=k % NSLOT, rule=lm.RULE_AVX2)
        kms = pli.last_kernel_ms
        if world > 1:  # global argmax over all slabs (every rank computes the same cell)
            res = xchg.finish(k % NSLOT, row_offsets, rows_global, lm.RULE_AVX2)
        return res, kms

    def run_steps(n, sampler=None):
        out = []
        for j in range(min(depth, n)):
            launch(j)
        for k in range(n):
            if k + depth < n:
                launch(k + depth)
            if sampler is not None and world == 1 and n >= 4 and k in (n // 4, n // 2, (3 * n) // 4):
                sampler.sample_nvml()  # the GPU is busy with scans k and k + 1 here
            out.append(finish(k))
            if sampler is not None:
                sampler.sample_kernel(pli.last_sm_clock_mhz)
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    run_steps(max(args.warmup, 1))
    stats = pli.last_stats
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = pli.kernel_launches
    if sampler is not None:
        sampler.sample_nvml()  # before the barrier: the stall an NVML query causes is over when timing starts
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    results = run_steps(args.steps, sampler)
    e1.record(stream)
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.sample_nvml()
    sync_all()
    clocks = sampler.result() if rank == 0 else None
    res = results[-1][0]
    kernel_ms = [r[1] for r in results]
    ms = e0.elapsed_time(e1)
    launches = pli.kernel_launches - launches0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.float64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_per_step = ms / args.steps
    value = total / (ms_per_step * 1e-3) / 1e9  # whole-job Gbp/s

    # ---- the same scan with the filter disabled (every cell gets the exact f32 sum), for reference
    exact = None
    if rank == 0 and not args.no_exact:
        pli.set_prune(0)
        ems = []
        for _ in range(3):
            r_ex = seq.score_argmax(pssm, rule=lm.RULE_AVX2)
            ems.append(pli.last_kernel_ms)
        pli.set_prune(1)
        if world == 1:
            assert (r_ex["position"], r_ex["score"]) == (res["position"], res["score"]), (r_ex, res)
        ex_ms = float(np.median(ems[1:]))
        exact = {"value": positions / (ex_ms * 1e-3) / 1e9, "unit": "Gbp/s (per GPU)", "kernel_ms": ex_ms,
                 "kernel": "lm::score_kernel<16, MODE_ARGMAX, f32> (LM_OPT_PRUNE=0: exact f32 sum at every cell)",
                 "same_result": True}
    if world > 1:
        dist.barrier()

    # ---- end to end through the C ABI with HOST buffers (rank-local slab)
    e2e = None
    if not args.no_e2e:
        host = torch.empty(((r1 - r0) + WIDTH - 1, 32), dtype=torch.uint8, pin_memory=True)
        dl = seq.download()
        host.numpy()[:] = dl
        del dl
        harr = host.numpy()
        seq.free()  # the resident copy is not used by the e2e path

        def e2e_step():
            s = pli.sequence_from_striped(harr, total, WIDTH - 1)   # H2D of this step's input
            r = s.score_argmax(pssm, rule=lm.RULE_AVX2)              # kernel + D2H of the result
            s.free()
            return r

        for _ in range(min(args.warmup, 2)):
            e2e_step()
        sync_all()
        n_e2e = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            r_e2e = e2e_step()
        sync_all()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": total / (dt / n_e2e) / 1e9, "unit": "Gbp/s",
               "h2d_bytes_per_step": int(harr.nbytes) * world + int(pssm.nbytes) * world,
               "d2h_bytes_per_step": 272 * world,
               "steps": n_e2e, "api": "lm_seq_from_striped(pinned host StripedSequence) + lm_seq_score_argmax_f32"}
        if world == 1:
            assert r_e2e["position"] == res["position"], (r_e2e, res)

    if rank == 0:
        peak, how = measured_peaks()
        k_ms = float(np.mean(kernel_ms))
        achieved = positions * 1.0 / (k_ms * 1e-3) / 1e9  # GB/s of algorithmic bytes, per GPU
        line = {
            "metric": METRIC, "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"width-20 DNA PSSM (JASPAR MA0138.3 REST) over {per_gpu / 1e9:g} Gbp synthetic DNA "
                            f"per GPU, fused score+argmax (configs[1]; x{world} GPUs row-sharded)",
                "bp_per_gpu": per_gpu, "total_bp": total, "seed": SEED,
                "l2": "input per GPU (5 GB) exceeds L2 (126 MB); no flush needed",
                "method": "LM_OPT_PRUNE=auto: integer upper bound for every cell (PRMT byte-lane look-ups, 4 cells per "
                          "instruction), exact sequential f32 sum only where the bound reaches the running maximum; "
                          "result bit-identical to scoring every cell (checked in-run against exact_everywhere)",
                "parallelism": f"row-sharded x{world}, NCCL all-gather of the 272-byte winner blocks (device-resident)" if world > 1 else "1 GPU",
                "result": {"position": res["position"], "score": res["score"]},
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(per_gpu, stats["pruned"]),
                         "traffic_source": "profiles/ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum of one "
                                           "launch, ncu --set full)", "peak_source": how,
                         "kernel": (f"lm::prune_kernel<{stats['filter_rows']}, {'paired' if stats.get('paired') else 'single'}, argmax> (exact filter + recheck; "
                                    "+ lm::score_kernel on the first 262144 rows)") if stats["pruned"]
                         else "lm::score_kernel<16, MODE_ARGMAX, f32>",
                         "kernel_ms": k_ms, "algorithmic_bytes_per_launch": positions,
                         "filter": stats},
            "exact_everywhere": exact,
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e,
        }
        if world == 1 and not args.no_cpu:
            cb, _ = cpu_baseline(pssm)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--bp-per-gpu", type=int, default=PER_GPU_BP)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-exact", action="store_true")
    ap.add_argument("--depth", type=int, default=2, choices=[1, 2, 3], help="scans enqueued ahead of the host")
    ap.add_argument("--reserve-sms", type=int, default=0, help="N > 1: SMs the scan grid leaves free for the NCCL kernel")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
