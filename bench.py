#!/usr/bin/env python
"""Benchmark of the sourmash hot path (FracMinHash sketch + all-vs-all intersection -> ANI).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|config3|config4|config5|tiny]
    python bench.py --impl reference ...        # the CPU arm (oracle port on the host cores)

One *step* = one pass of the hot path over one batch of synthetic genomes of the named BASELINE.json
configuration: kernel K1 (hash + per-genome sort/dedup) over every genome, [all-gather of the
sketches when N > 1], kernel K2 over every unordered genome pair, and the ANI kernel.

* ``value``  whole-job genome pairs/s with the 2-bit packed genomes already resident in HBM.
* ``e2e``    the same metric through the public engine API from HOST buffers: the ASCII genomes are packed
             (2 bits per base; validity masks only for tiles that hold an invalid base) by the library's
             host threads into pinned memory and copied host->device chunk by chunk, chunks from the tail of
             the stream cross as plain ASCII while the link would otherwise idle and are packed on the GPU,
             K1 hashes the chunks already there, K2 intersects, and the two float64 ANI matrices are copied
             back -- all inside the timed region.
* ``roofline``       dominant kernel of the step, algorithmic bytes / CUDA-event time vs measured HBM peak.
* ``cpu_baseline``   the oracle (CPU port of the same algorithm) timed on the host cores (rank 0).

Timing: CUDA events on the launching stream around every step (max over ranks), >= 3 warm-up steps,
L2 flushed between steps by writing a 256 MiB buffer (outside the event pairs); the whole timed loop is
bracketed by barrier + synchronize.  Multi-GPU (``torchrun``): strong scaling, genomes sliced across
ranks for K1, one NCCL all-gather of the sketch rows, K2 work items dealt round-robin to ranks.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

SEED = 20261017
METRIC = "all-vs-all genome pairs/sec + sketch Gbp/s (whole-job genome pairs/s; sketch Gbp/s in extra keys)"
UNIT = "genome pairs/s"

WORKLOADS = {
    # name: (n_genomes, genome length, k, scaled, description from BASELINE.json configs[])
    "tiny": (16, 200_000, 31, 1000, "16 synthetic 0.2 Mb genomes (self-test only)"),
    "config2": (100, 5_000_000, 31, 1000, "configs[1]: 100 synthetic 5 Mb genomes, k=31, scaled=1000"),
    "config3": (1000, 5_000_000, 31, 1000, "configs[2]: 1,000 synthetic 5 Mb genomes, k=31, scaled=1000"),
    "config4": (10000, 5_000_000, 31, 1000, "configs[3]: 10,000 synthetic 5 Mb genomes, k=31, scaled=1000"),
    "config5": (2000, 5_000_000, 31, 100, "configs[4]: 2,000 synthetic 5 Mb genomes, k=31, scaled=100"),
}


# translation unit + headers each captured kernel is compiled from (include graph of csrc/): an ncu capture
# stays evidence for as long as THESE files are unchanged, whatever happens to the other kernels' sources
KERNEL_SOURCES = {
    "k1": ("sketch.cu", "kmer_hash.cuh", "pack.cuh", "common.cuh"),
    "k2_index": ("index.cu", "common.cuh"),
    "k2_probe": ("pairwise.cu", "common.cuh"),
}


def kernel_source_sha(kernel: str | None = None) -> str:
    """sha256 over the CUDA sources a kernel is built from (``kernel`` in KERNEL_SOURCES: its translation unit,
    the headers it includes and the preprocessor lines of include/panib200.h -- the only part of the C header
    that reaches device code; None = every CUDA source of the library and the whole header): ties an ncu
    capture to the code it measured."""
    import hashlib

    h = hashlib.sha256()
    csrc = ROOT / "pyani_plus_b200" / "csrc"
    header = ROOT / "include" / "panib200.h"
    if kernel is None:
        for f in sorted([*csrc.glob("*.cu"), *csrc.glob("*.cuh"), header]):
            h.update(f.name.encode())
            h.update(f.read_bytes())
    else:
        for name in sorted(KERNEL_SOURCES[kernel]):
            h.update(name.encode())
            h.update((csrc / name).read_bytes())
        h.update(b"".join(ln for ln in header.read_bytes().splitlines(keepends=True) if ln.lstrip().startswith(b"#")))
    return h.hexdigest()[:16]


def ncu_capture(workload: str, kernel: str) -> dict | None:
    """Per-launch ncu numbers of ``kernel`` ("k1", "k2_probe", "k2_index") at ``workload`` from the newest
    profiles/ncu_r*.json (written by tools/ncu_to_json.py from an ``ncu --set full`` capture) -- but only if
    that capture was taken from the sources this tree builds the kernel from; otherwise None (stale = not
    evidence)."""
    files = sorted((ROOT / "profiles").glob("ncu_r*.json"))
    if not files:
        return None
    try:
        data = json.loads(files[-1].read_text())
    except ValueError:
        return None
    cap = data.get("captures", {}).get(workload, {}).get(kernel)
    if cap is None:
        return None
    if "source_sha" in cap:
        fresh = cap["source_sha"] == kernel_source_sha(kernel)
    else:  # older files: one hash over the whole library
        fresh = data.get("source_sha") == kernel_source_sha()
    return dict(cap, file=f"profiles/{files[-1].name}") if fresh else None


def config_dict(workload: str, n_gpus: int) -> dict:
    """The ``config`` object of the JSON line: a function of (workload, N) only, so that the GPU arm and the
    reference arm print the same one."""
    n, length, k, scaled, desc = WORKLOADS[workload]
    return {"workload": desc, "n_genomes": n, "genome_bp": length, "k": k, "scaled": scaled,
            "pairs": n * (n - 1) // 2, "seed": SEED, "l2": "flushed between steps (256 MiB write)",
            "parallelism": (f"genomes sliced over {n_gpus} ranks for K1, one exchange of the sketches over "
                            "NVLink, K2 work sharded by rank") if n_gpus > 1 else "single GPU"}


def expected_checksum(workload: str) -> dict | None:
    """Oracle-pinned result checksum of a workload (tools/oracle_checksums.py), or None if not pinned."""
    p = ROOT / "tests" / "golden" / "workload_checksums.json"
    if not p.is_file():
        return None
    e = json.loads(p.read_text()).get(workload)
    return {k_: e[k_] for k_ in ("ov_weighted_sum", "hash_sum", "sketch_total")} if e else None


def checksum_weights(rows, cols):
    """Position weights of the count-matrix checksum (numpy int64 arrays or torch int64 tensors)."""
    return (rows[:, None] * 1000003 + cols[None, :] * 7919 + 1) % 2147483647


def hbm_peak() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except (ValueError, KeyError):
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML from a thread beside the timed region."""

    REASONS = {
        "hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
        "hw_thermal_slowdown": 0x40, "hw_power_brake_slowdown": 0x80,
    }

    def __init__(self, gpu_index: int, period_s: float = 0.002) -> None:
        import threading

        self.period = period_s
        self.samples: list[tuple[int, int]] = []
        self.sm_max = None
        self.err = None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES remapping when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                try:
                    phys = int(vis.split(",")[gpu_index])
                except (ValueError, IndexError):
                    phys = gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv = None
            self.err = f"NVML unavailable: {e}"

    def _run(self) -> None:
        nv = self.nv
        while not self._stop.is_set():
            try:
                clk = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    rs = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((clk, rs))
            except Exception as e:  # noqa: BLE001
                self.err = str(e)
                return
            time.sleep(self.period)

    def start(self) -> None:
        if self.nv is not None:
            self._thread.start()

    def stop(self) -> dict:
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": [self.err or "no NVML"]}
        self._stop.set()
        self._thread.join(timeout=2)
        clocks = [c for c, _ in self.samples]
        reasons = set()
        for _, rs in self.samples:
            for name, bit in self.REASONS.items():
                if rs & bit:
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(clocks)) if clocks else None,
            "sm_max_mhz": self.sm_max,
            "samples": len(clocks),
            "reasons": sorted(reasons),
        }


def integer_pipe_view(bases: int, k1_ms: float, clocks: dict | None, cap: dict | None) -> dict:
    """What actually bounds K1: issue slots and the two half-rate integer pipes.

    Warp instructions per launch and the pipe-busy percentages come from the ncu capture of THESE kernel
    sources (``ncu_capture``; null when the committed capture is of other sources); time and SM clock are
    measured in this run.  A B200 SM issues at most one warp instruction per cycle on each of its 4
    schedulers; the ALU and FMA-heavy pipes take one per two cycles each.
    """
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    slots = k1_ms * 1e-3 * mhz * 1e6 * 148 * 4
    out = {"sm_mhz_used": mhz, "issue_slots_in_launch": slots, "warp_inst_per_kmer": None, "issue_slot_frac": None,
           "ncu_pipe_busy": None, "source": None}
    if cap and cap.get("inst_executed") and cap.get("bases"):
        per_kmer = cap["inst_executed"] / (cap["bases"] / 32.0)
        out.update({"warp_inst_per_kmer": per_kmer, "issue_slot_frac": bases / 32.0 * per_kmer / slots,
                    "ncu_pipe_busy": cap.get("pipe_busy"), "source": cap.get("file")})
    return out


# ======================================================================================= CPU arm
def cpu_measure(workload: str, steps: int, warmup: int, budget_s: float = 25.0) -> dict:
    """Time the oracle port (all host threads) on the workload, or on a bounded sample of it.

    The whole workload is timed when ``steps + warmup`` passes of it fit ``budget_s`` (estimated from a
    one-genome probe); otherwise the largest genome subset that fits, and the whole-job time is
    extrapolated: sketching scales linearly in the number of genomes, intersection in the number of pairs.
    """
    from oracle import oracle

    n, length, k, scaled, _ = WORKLOADS[workload]
    threads = oracle.num_threads()
    n_pairs = n * (n - 1) // 2
    lib = oracle.lib()
    cap = int(length / scaled * 1.5) + 256
    # probe: one genome per thread, to size the sample from this machine's speed
    np_ = min(n, threads)
    probe = np.empty((np_, length), dtype=np.uint8)
    for g in range(np_):
        probe[g] = np.frombuffer(oracle.synth_genome(SEED, g, length), dtype=np.uint8)
    out = np.zeros((np_, cap), dtype=np.uint64)
    counts = np.zeros(np_, dtype=np.int64)
    t0 = time.perf_counter()
    lib.oracle_sketch_batch(probe.ctypes.data, np_, length, k, oracle.max_hash(scaled),
                            out.ctypes.data_as(oracle.c_u64p), cap, counts.ctypes.data_as(oracle.c_i64p))
    per_genome_s = (time.perf_counter() - t0) / np_  # wall seconds per genome with all threads busy
    per_step = budget_s / max(1, steps + warmup)
    pair_s = 2.5e-6 * (1000.0 / scaled) * 16 / max(1, threads)  # rough merge cost per pair, only sizes the sample
    ns = n
    if n * per_genome_s + n_pairs * pair_s > per_step:
        ns = int(max(min(n, threads), min(n, per_step / (per_genome_s + pair_s * n / 2))))
        while ns > threads and ns * per_genome_s + ns * (ns - 1) / 2 * pair_s > per_step:
            ns = int(ns * 0.9)
    seqs = np.empty((ns, length), dtype=np.uint8)
    seqs[:np_] = probe[:min(np_, ns)] if ns >= np_ else probe[:ns]
    for g in range(min(np_, ns), ns):
        seqs[g] = np.frombuffer(oracle.synth_genome(SEED, g, length), dtype=np.uint8)
    del probe
    out = np.zeros((ns, cap), dtype=np.uint64)
    counts = np.zeros(ns, dtype=np.int64)
    t_sk, t_ix = [], []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        lib.oracle_sketch_batch(seqs.ctypes.data, ns, length, k, oracle.max_hash(scaled),
                                out.ctypes.data_as(oracle.c_u64p), cap, counts.ctypes.data_as(oracle.c_i64p))
        t1 = time.perf_counter()
        ov = oracle.intersect_all(out, counts)
        # ANI finalisation is part of the path (one row's worth: the Python loop is not what is measured)
        for i in range(min(ns, 64)):
            oracle.pair_row(int(ov[0, i]), int(counts[0]), int(counts[i]), k)
        t2 = time.perf_counter()
        if it >= warmup:
            t_sk.append(t1 - t0)
            t_ix.append(t2 - t1)
    sk = float(np.mean(t_sk))
    ix = float(np.mean(t_ix))
    s_pairs = ns * (ns - 1) // 2
    full = ns == n
    full_s = sk + ix if full else sk * (n / ns) + ix * (n_pairs / max(1, s_pairs))
    sample = (f"the whole workload per step: {n} genomes sketched ({sk:.3f} s) and all {n_pairs} pairs "
              f"intersected ({ix:.4f} s)" if full else
              f"{ns} of {n} genomes sketched ({sk:.3f} s) and their {s_pairs} pairs intersected ({ix:.4f} s) "
              f"per step, extrapolated linearly in genomes / pairs to the full workload")
    return {
        "value": n_pairs / full_s, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
        "sketch_gbp_s": ns * length / sk / 1e9,
        "pairs_per_s_intersect": s_pairs / ix if ix > 0 else None,
        "step_s": sk + ix, "full_workload_s": full_s, "sampled_genomes": ns, "extrapolated": not full,
    }


def run_reference(args: argparse.Namespace) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun sets OMP_NUM_THREADS=1 for its workers; this arm runs on rank 0 alone and must use
        # every host thread it can (set before the OpenMP runtime of the oracle library starts)
        os.environ["OMP_NUM_THREADS"] = str(len(os.sched_getaffinity(0)))
    base = cpu_measure(args.workload, args.steps, args.warmup, budget_s=args.reference_budget)
    line = {
        "impl": "reference",
        "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": base["full_workload_s"] * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_dict(args.workload, args.gpus),
        "note": "pyani-plus's real CPU path (sourmash + branchwater, Rust) is not installable here; this arm "
                "times the oracle CPU port of the same algorithm with OpenMP on all host threads",
        "cpu_baseline": {k_: base[k_] for k_ in ("value", "unit", "cores", "kind", "sample")},
        "sketch_gbp_s": base["sketch_gbp_s"],
        "pairs_per_s_intersect": base["pairs_per_s_intersect"],
        "measured_step_s": base["step_s"], "extrapolated": base["extrapolated"],
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ======================================================================================= GPU arm
def run_gpu(args: argparse.Namespace) -> None:  # noqa: PLR0915
    import torch
    import torch.distributed as dist

    from pyani_plus_b200 import engine, multi_gpu
    from pyani_plus_b200 import stream as pstream

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    # stdout carries exactly ONE line (the JSON): whatever libraries print there while the job runs
    # (NCCL's version banner, for one) is sent to stderr; the descriptor is restored for the result
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    from pyani_plus_b200 import run as run_mod

    ctx = run_mod.DistContext.from_env()  # torchrun: one process per GPU, NCCL
    world, rank = ctx.world, ctx.rank
    eng = engine.Engine(local_rank)
    dev = eng.device

    n, length, k, scaled, desc = WORKLOADS[args.workload]
    n_pairs = n * (n - 1) // 2
    # genomes sketched by each rank (last ranks may hold empty dummies so that the gather is fixed-size)
    g0, g1, per_rank = multi_gpu.slice_for_rank(n, rank, world)
    n_local = g1 - g0

    # ---- inputs: this rank's slice of the genomes, generated on the device, then packed (resident)
    # dummy genomes (length 0) pad the slice so that every rank gathers the same shape
    tiles_per = length // pstream.TILE + 1
    tile_off = np.zeros(per_rank + 1, dtype=np.int64)
    for i in range(per_rank):
        tile_off[i + 1] = tile_off[i] + (tiles_per if i < n_local else 1)
    plan = run_mod.agree_row_stride(eng, tile_off, scaled, ctx)  # one row stride on every rank
    bufs = eng.alloc_stream_buffers(plan, ascii_too=False)
    tab = eng.alloc_table(plan)
    # e2e needs the whole ASCII stream in pinned host memory; skipped (null) for streams > 12 GB
    do_e2e = plan.n_bases <= (12 << 30) and not args.no_e2e
    # what a FASTA reader hands over: the ASCII stream in page-locked host memory
    h_ascii = torch.empty(plan.n_bases, dtype=torch.uint8, pin_memory=True) if do_e2e else None
    batch = 256  # genomes generated + packed per pass, so that the ASCII form is never fully resident
    for b0 in range(0, per_rank, batch):
        b1 = min(per_rank, b0 + batch)
        t0, t1 = int(tile_off[b0]), int(tile_off[b1]) + (1 if b1 == per_rank else 0)  # + the trailing pad tile
        d_ascii = torch.full(((t1 - t0) * pstream.TILE,), pstream.PAD, dtype=torch.uint8, device=dev)
        real = max(0, min(b1, n_local) - b0)
        if real:
            gen, _ = eng.synth_ascii_stream(SEED, g0 + b0, real, length)
            d_ascii[: real * tiles_per * pstream.TILE] = gen[: real * tiles_per * pstream.TILE]
            del gen
        engine._check(eng.lib.panib_pack_ascii(  # noqa: SLF001
            d_ascii.data_ptr(), d_ascii.numel(), bufs["packed"].data_ptr() + t0 * pstream.TILE // 4,
            bufs["mask"].data_ptr() + t0 * pstream.TILE // 8, eng._stream()))  # noqa: SLF001
        if do_e2e:
            h_ascii[t0 * pstream.TILE: t1 * pstream.TILE].copy_(d_ascii)
        torch.cuda.synchronize()
        del d_ascii
    if do_e2e:  # pinned staging of the packed form: the host threads pack into it inside every e2e step
        bufs["h_packed"] = torch.empty(plan.n_bases // 16, dtype=torch.int32, pin_memory=True)
        bufs["h_mask"] = torch.empty(plan.n_bases // 32, dtype=torch.int32, pin_memory=True)
        eng.add_ingest_scratch(plan, bufs)

    n_rows = per_rank * world
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    result = {}

    # the step is the package's (run.make_step): size hint agreed over the ranks; several GPUs: finalize fused
    # with the all-gather through peer-mapped memory (an error if that cannot be set up, unless --nccl-gather)
    stepper, exchange = run_mod.make_step(
        eng, plan, bufs, tab, k, ctx, h_ascii=h_ascii, k2_method=args.k2, nccl_gather=args.nccl_gather,
        host_threads=args.host_threads or max(1, len(os.sched_getaffinity(0)) // max(1, world)))
    size_hint = stepper.size_hint
    # the step as ONE CUDA graph launch (all ranks must agree, the gather has barriers inside)
    graphed: dict = {}
    per_step_launches: dict = {}
    if not args.no_graph:
        for from_host in ((False, True) if do_e2e else (False,)):
            l0 = eng.launch_count()
            ok = stepper.capture(from_host=from_host, to_host=from_host)
            per_step_launches[from_host] = (eng.launch_count() - l0) // 2  # warm-up run + capture
            ok_t = torch.tensor([1 if ok else 0], device=dev)
            if world > 1:
                dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
            graphed[from_host] = bool(int(ok_t.item()))

    def step(from_host: bool, marks: list | None = None, *, graph: bool = False) -> None:
        """One pass of the hot path (product API: pipeline.SourmashStep), status checked at the end."""
        if graph:
            stepper.replay(from_host=from_host, to_host=from_host)
        else:
            stepper.enqueue(from_host=from_host, marks=marks, to_host=from_host)
        stepper.finish()
        out = stepper.out
        if from_host:
            result["identity"], result["cov_query"] = out["identity_host"], out["cov_query_host"]
            result["counts"] = out["counts_host"]
        else:
            result["ov"], result["table"] = out["ov"], out["table"]

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(from_host: bool, steps: int, warmup: int, *, graph: bool = False) -> dict:
        for _ in range(warmup):
            step(from_host, graph=graph)
            flush.fill_(1)
        # the clock sampler starts BEFORE the barrier: starting it costs rank 0 a millisecond or two, which
        # every other rank would otherwise wait for inside the first timed step's gather
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        barrier()
        launches0 = eng.launch_count()
        tot, t_k1, t_gather, t_k2 = [], [], [], []
        wall0 = time.perf_counter()
        for _ in range(steps):
            e0, e1 = ev(), ev()
            marks = [ev(), ev(), ev()]
            e0.record()
            step(from_host, None if graph else marks, graph=graph)
            if graph:
                for m in marks:  # no stage marks inside a graph launch
                    m.record()
            e1.record()
            e1.synchronize()
            tot.append(e0.elapsed_time(e1))
            t_k1.append(e0.elapsed_time(marks[0]))
            t_gather.append(marks[0].elapsed_time(marks[1]))
            t_k2.append(marks[1].elapsed_time(marks[2]))
            flush.fill_(1)  # L2 flush, outside the event pair
        barrier()
        if os.environ.get("PANIB_BENCH_DEBUG"):
            print(f"[rank {rank}] from_host={from_host} step ms {[round(x, 3) for x in tot]} "
                  f"k1 {[round(x, 3) for x in t_k1]} gather {[round(x, 3) for x in t_gather]}", file=sys.stderr)
        wall = time.perf_counter() - wall0
        clocks = sampler.stop() if rank == 0 else None
        launches = eng.launch_count() - launches0 + (steps * per_step_launches[from_host] if graph else 0)
        ing = eng.ingest_last() if from_host else {"h2d_bytes": 0, "chunks": 0, "chunks_as_ascii": 0, "dirty_tiles": 0,
                                                   "ring_bytes": 0}
        stats = torch.tensor([sum(tot), sum(t_k1), sum(t_gather), sum(t_k2), float(launches),
                              float(ing["h2d_bytes"]), float(ing["chunks"]), float(ing["chunks_as_ascii"]),
                              float(ing["dirty_tiles"]), float(ing["ring_bytes"] > 0)], dtype=torch.float64, device=dev)
        if world > 1:
            mx = stats.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = stats.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            stats = torch.cat([mx[:4], sm[4:]])
        s = stats.cpu().tolist()
        return {"ms": s[0] / steps, "k1_ms": s[1] / steps, "gather_ms": s[2] / steps, "k2_ms": s[3] / steps,
                "launches": int(s[4]), "wall_s": wall, "clocks": clocks,
                "ingest": {"h2d_bytes": int(s[5]), "chunks": int(s[6]), "chunks_as_ascii": int(s[7]),
                           "dirty_tiles": int(s[8]), "ranks_using_the_cached_ring": int(s[9])}}

    eager_t = timed_loop(False, args.steps, args.warmup)  # stage breakdown (events between the stages)
    dev_t = timed_loop(False, args.steps, args.warmup, graph=True) if graphed.get(False) else eager_t
    for key in ("k1_ms", "gather_ms", "k2_ms"):
        dev_t[key] = eager_t[key]

    # kernel-only timing of the dominant kernel (K1 hash) and of K2, alone on the stream
    def time_kernel(fn, reps: int) -> float:
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            a, b = ev(), ev()
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))

    sk_args = eng._sketch_args(plan, bufs, tab, k, 42)  # noqa: SLF001
    hash_args = sk_args[:12] + (sk_args[13], sk_args[14], sk_args[15])

    def hash_only() -> None:
        engine._check(eng.lib.panib_sketch_hash_only(*hash_args))  # noqa: SLF001

    def finalize_only() -> None:
        engine._check(eng.lib.panib_sketch_finalize(tab["table"].data_ptr(), plan.row_stride, plan.n_genomes,  # noqa: SLF001
                                                    plan.d_nb.data_ptr(), tab["counts"].data_ptr(),
                                                    tab["flags"].data_ptr(), 0, eng._stream()))  # noqa: SLF001

    reps = max(3, min(args.steps, 10))
    k1_hash_ms = time_kernel(hash_only, reps)
    finalize_only()
    step(False)
    table = result["table"]
    k2_ms = time_kernel(lambda: eng.intersect(table, rank=rank, world=world, max_count=size_hint, check=False,
                                              method=stepper.k2_method), reps)
    assert eng.check_status() == 0
    counts_host = table.counts.cpu().numpy()[multi_gpu.real_rows(n, world)].astype(np.int64)

    # ---- result checksums (outside every timed region): must be identical for every --gpus N
    real = torch.from_numpy(multi_gpu.real_rows(n, world)).to(dev)
    idx = torch.arange(n, device=dev, dtype=torch.int64)
    ov_ck = torch.zeros(1, dtype=torch.int64, device=dev)
    for r0 in range(0, n, 1024):  # row blocks keep the int64 temporaries small at 10,000 genomes
        blk = result["ov"][real[r0: r0 + 1024]][:, real].to(torch.int64)
        w = checksum_weights(idx[r0: r0 + 1024], idx)
        ov_ck += (blk * w).sum()
    if world > 1:
        dist.all_reduce(ov_ck, op=dist.ReduceOp.SUM)  # the ranks hold disjoint parts of the matrix
    rows_real = table.rows[real]
    valid = torch.arange(rows_real.shape[1], device=dev)[None, :] < table.counts[real][:, None]
    hash_ck = int((rows_real * valid).sum().item())  # wrapping int64 sum of every sketch hash
    checksum = {"ov_weighted_sum": int(ov_ck.item()), "hash_sum": hash_ck, "sketch_total": int(counts_host.sum())}
    del rows_real, valid
    # parity gate: the oracle's checksum of the COMPLETE workload (tests/golden/workload_checksums.json)
    expected = expected_checksum(args.workload)
    parity = {"expected": expected, "ok": (checksum == expected) if expected else None,
              "source": "tests/golden/workload_checksums.json (tools/oracle_checksums.py)" if expected else
                        "workload not pinned"}

    e2e_t = None
    if do_e2e:
        e2e_eager = timed_loop(True, max(2, args.steps // 2), 3)
        e2e_t = timed_loop(True, max(2, args.steps // 2), 3, graph=True) if graphed.get(True) else e2e_eager
        for key in ("k1_ms", "gather_ms", "k2_ms"):
            e2e_t[key] = e2e_eager[key]

    if rank != 0:
        ctx.close()
        if parity["ok"] is False:
            sys.exit(3)
        return

    # ---- derived numbers
    peak, peak_src = hbm_peak()
    local_bases = n_local * length
    k1_bytes = local_bases * (0.25 + 0.125 + 8.0 / scaled)  # 2-bit stream + validity mask + hash writes
    k1_gbs = k1_bytes / (k1_hash_ms * 1e-3) / 1e9
    tot_cnt = int(counts_host.sum())
    # K2 algorithmic bytes: sum over unordered pairs of 8(|A|+|B|) + 4 = 8 (n-1) sum|A| + 4 pairs
    k2_bytes = (8.0 * (n - 1) * tot_cnt + 4.0 * n_pairs) / world
    k2_gbs = k2_bytes / (k2_ms * 1e-3) / 1e9
    value = n_pairs / (dev_t["ms"] * 1e-3)
    e2e_value = n_pairs / (e2e_t["ms"] * 1e-3) if e2e_t else None
    dominant_is_k1 = dev_t["k1_ms"] >= dev_t["k2_ms"]
    cap_k1 = ncu_capture(args.workload, "k1")
    cap_k2 = ncu_capture(args.workload, "k2_" + stepper.k2_method)
    roof_k1 = {"kernel": "sketch_hash_kernel<31> (K1)", "bound": "hbm", "achieved": k1_gbs, "peak": peak,
               "unit": "GB/s", "frac": k1_gbs / peak,
               "traffic": (cap_k1 or {}).get("dram_bytes") if world == 1 else None,
               "algorithmic_bytes": k1_bytes, "peak_source": peak_src,
               "ms_per_launch": k1_hash_ms, "bytes_per_bp": 0.25 + 0.125 + 8.0 / scaled,
               "note": "integer-ALU bound by construction (MurmurHash3 over 31 ASCII bytes per base); see "
                       "DESIGN.md and profiles/ for pipe utilisation",
               "integer_pipes": integer_pipe_view(local_bases, k1_hash_ms, dev_t["clocks"], cap_k1)}
    if stepper.k2_method == "probe":
        roof_k2 = {"kernel": "intersect_kernel (K2, probing form)", "bound": "hbm", "achieved": k2_gbs, "peak": peak,
                   "unit": "GB/s", "frac": k2_gbs / peak,
                   "traffic": (cap_k2 or {}).get("dram_bytes") if world == 1 else None,
                   "algorithmic_bytes": k2_bytes, "peak_source": peak_src, "ms_per_launch": k2_ms,
                   "bytes_per_pair": k2_bytes * world / max(1, n_pairs),
                   "note": "algorithmic bytes 8(|A|+|B|)+4 per pair (SURVEY 8d); staged queries and L2-resident "
                           "columns make DRAM traffic far smaller, so frac can exceed 1"}
    else:
        # The inverted-index form never touches what SURVEY 8d's per-pair definition counts (it reads every sketch
        # ONCE), so it gets the byte model of its own kernels: what each of them must move once, per rank.
        est = eng.last_intersect_estimates or {}
        own = tot_cnt if world == 1 else min(tot_cnt, 2 * (tot_cnt // world) + 4096)
        slots = 1024
        while slots < 2 * own:
            slots <<= 1
        cols = int(est.get("frequent_hashes", 0))
        rare_pairs = int(est.get("rare_pairs", 0))
        t64 = (n + 63) // 64
        idx_bytes = (16.0 * slots                       # hash table cleared: key 8 + count 4 + aux 4 per slot
                     + 8.0 * tot_cnt                    # insert: every rank reads every entry's hash
                     + (20.0 + 12.0 + 16.0 + 16.0) * tot_cnt / world  # insert (slot CAS + count + tag write), classify,
                                                                      # emit, sparse: tags / counts read per owned entry
                     + 8.0 * rare_pairs                 # one 4-byte atomic read-modify-write per rare pair
                     + t64 * (t64 + 1) / 2 * 2 * 64 * ((cols + 31) // 32) * 4.0  # bit-matrix rows per 64x64 tile
                     + 8.0 * n * n)                     # mirror: count matrix read and written
        idx_gbs = idx_bytes / (k2_ms * 1e-3) / 1e9
        roof_k2 = {"kernel": "index_* kernels (K2, inverted-index form: hash-table group-by + AND/POPC bit matrix "
                             "+ rare-pair adds)", "bound": "hbm", "achieved": idx_gbs, "peak": peak, "unit": "GB/s",
                   "frac": idx_gbs / peak,
                   "traffic": (cap_k2 or {}).get("dram_bytes") if world == 1 else None,
                   "algorithmic_bytes": idx_bytes, "peak_source": peak_src, "ms_per_launch": k2_ms,
                   "bytes_per_pair": idx_bytes * world / max(1, n_pairs),
                   "per_pair_definition": {"algorithmic_bytes": k2_bytes, "achieved": k2_gbs, "frac": k2_gbs / peak},
                   "model": {"table_slots": slots, "entries": tot_cnt, "bit_matrix_columns": cols,
                             "rare_pairs": rare_pairs},
                   "note": "byte model of the form's own kernels (DESIGN.md 4): table clear 16 B/slot, 8 B/entry read "
                           "by every rank, 64 B/owned entry of slot / tag / list accesses over insert, classify, emit "
                           "and sparse, 8 B per rare pair, bit-matrix rows per 64x64 tile, N^2 x 8 B mirror; "
                           "per_pair_definition keeps SURVEY 8d's 8(|A|+|B|)+4, which this form does not move "
                           "(frac >> 1 there)"}
    # the CPU leg runs at N=1 only (under torchrun the workers are pinned to one OpenMP thread)
    cpu = cpu_measure(args.workload, 2, 1) if (world == 1 and not args.no_cpu_baseline) else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_t["ms"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_dict(args.workload, world),
        "exchange": exchange,
        "sketch_gbp_s": n * length / (dev_t["k1_ms"] * 1e-3) / 1e9,
        "pairs_per_s_k2": n_pairs / (dev_t["k2_ms"] * 1e-3),
        "stage_ms": {"k1_sketch": dev_t["k1_ms"], "allgather": dev_t["gather_ms"], "k2_intersect": dev_t["k2_ms"],
                     "k1_hash_kernel_alone": k1_hash_ms, "k2_alone": k2_ms},
        "roofline": roof_k1 if dominant_is_k1 else roof_k2,
        "roofline_k1": roof_k1, "roofline_k2": roof_k2,
        "cpu_baseline": ({k_: cpu[k_] for k_ in ("value", "unit", "cores", "kind", "sample")} if cpu else None),
        "cpu_baseline_detail": ({"sketch_gbp_s": cpu["sketch_gbp_s"],
                                 "pairs_per_s_intersect": cpu["pairs_per_s_intersect"]} if cpu else None),
        "e2e": ({"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_t["ms"],
                 "h2d_bytes_per_step": e2e_t["ingest"]["h2d_bytes"],
                 "host_ascii_bytes_per_step": int(plan.n_bases) * world,
                 "ingest": {**e2e_t["ingest"],
                            "what": "counted by the library in the last timed step, summed over ranks: packed chunks "
                                    "(0.25 B/base) + masks of the tiles holding invalid bases + chunks sent as ASCII "
                                    "(1 B/base) from the tail of the stream while the link would otherwise idle"},
                 "host_pack": {"threads_per_rank": min(stepper.host_threads, int(eng.lib.panib_host_threads())),
                               "what": "ASCII -> 2-bit + dirty-tile masks on the host threads (AVX-512/AVX2), "
                                       "inside the timed step, pipelined with the copies and K1"},
                 "d2h_bytes_per_step": int(2 * n_rows * n_rows * 8 + n_rows * 4) * world,
                 "stage_ms": {"h2d_pack_k1": e2e_t["k1_ms"], "allgather": e2e_t["gather_ms"],
                              "k2_intersect": e2e_t["k2_ms"]}} if e2e_t else
                {"value": None, "unit": UNIT, "skipped": "ASCII stream larger than the 12 GB host staging limit "
                                                         "of bench.py (or --no-e2e)"}),
        "gpu_launches": dev_t["launches"],
        "k2_method": {"used": stepper.k2_method, "requested": args.k2, "estimates": eng.last_intersect_estimates},
        "cuda_graph": {"device_step": bool(graphed.get(False)), "e2e_step": bool(graphed.get(True)),
                       "eager_ms_per_step": eager_t["ms"],
                       "eager_e2e_ms_per_step": e2e_eager["ms"] if do_e2e else None,
                       "note": "value / e2e time one graph launch per step when captured; stage_ms come "
                               "from an eager pass with events between the stages"},
        "clocks": dev_t["clocks"],
        "sketch_sizes": {"mean": float(counts_host.mean()), "max": int(counts_host.max())},
        "result_checksum": checksum,
        "parity": parity,
        "library": engine.library_version(),
    }
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    print(json.dumps(line), flush=True)
    ctx.close()
    if parity["ok"] is False:
        print(f"PARITY FAILURE: result checksum {checksum} != oracle {expected}", file=sys.stderr, flush=True)
        sys.exit(3)


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="config3")
    ap.add_argument("--reference-budget", type=float, default=240.0,
                    help="--impl reference: seconds of CPU work the whole run may take; the full workload is "
                         "timed every step when it fits, else a bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--host-threads", type=int, default=0,
                    help="host threads per rank that pack ASCII in the e2e step (default: the rank's share of the cores)")
    ap.add_argument("--k2", default="auto", choices=["auto", "probe", "index"],
                    help="pairwise kernel: probing form, inverted-index form, or chosen from the data")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly (no CUDA graph replay)")
    ap.add_argument("--nccl-gather", action="store_true",
                    help="multi-GPU: plain NCCL all-gather instead of the fused finalize + peer-memory scatter")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
