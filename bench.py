#!/usr/bin/env python
"""bench.py — headline benchmark of the grafted CuArray reduce / scan / sort path on B200.

Metric (BASELINE.json): HBM GB/s for mapreduce / scan; sort Gkeys/s.  The headline line is
quoted on BASELINE.json configs[1] — `sum(A; dims=1)`, `sum(A; dims=2)`, `maximum`, `count` on
a 16384 x 16384 Float32 / BFloat16 CuMatrix (count: Bool matrix) — one "step" = those seven
reductions once: 5.10 GB of algorithmic input bytes.  `value` = algorithmic bytes / device time
with the matrices resident in HBM; `e2e` = the same step through the public host API starting
from pinned HOST buffers (H2D of all three matrices and D2H of every result inside the timed
region).  `extras` carries the other north-star configurations measured in the same run
(sum / cumsum / sort! / sortperm on 2^30 elements), each with its own roofline fraction.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extras]

N > 1 (torchrun): every rank owns one 16384 x 16384 column block of a 16384 x (16384*N) matrix
(weak scaling).  dims=1 sums need no exchange; dims=2 sums allreduce the 16384-vector of row
sums; maximum / count allreduce a scalar (NCCL).  value = bytes of all ranks / max-over-ranks time.
After the headline measurement the same launch times the sharded configs[2..4] paths (`extras.sharded`: cumsum and
sort! / sortperm of 2^30 elements in total, mapreduce(abs2, +) / extrema over 2^30 BFloat16 per GPU); they are guarded
by a watchdog and an exception handler so that they can never cost the JSON line (`--no-extras` skips them).

`--impl reference` times the reference's CPU path (Base Julia semantics restated in
oracle/oracle_cpu.c — the reference itself is Julia and no julia binary exists in this image)
on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M = 16384                       # matrix side of configs[1]
F32_BYTES = M * M * 4
BF16_BYTES = M * M * 2
BOOL_BYTES = M * M
STEP_BYTES = 3 * F32_BYTES + 3 * BF16_BYTES + BOOL_BYTES    # 5.10 GB algorithmic input bytes per step
NOMINAL_GBS = 8000.0


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            sm, smax, reasons = [], [], set()
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1])); smax.append(float(f[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            if sm:
                out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                       "samples": len(sm)}
        except Exception:
            pass
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return out


# ============================================================================ reference arm
def run_reference(args, rank, world):
    """The reference's CPU path (Base Julia semantics, C restatement) on the host cores."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_cpu as oc
    # bounded sample: a 4096-column block of the 16384 x 16384 matrices (1/4 of the step's bytes)
    cols = 4096
    rng = np.random.default_rng(0xB200 + 1)
    A = np.asfortranarray(rng.uniform(-1, 1, size=(M, cols)).astype(np.float32))
    H = np.asfortranarray((A.view(np.uint32) >> 16).astype(np.uint16))      # bf16 bits (truncation is fine for timing)
    Bm = np.asfortranarray(rng.random((M, cols)) < 0.5)
    sample_bytes = 3 * A.nbytes + 3 * H.nbytes + Bm.nbytes

    def step():
        oc.sum_dims(A, 1); oc.sum_dims(A, 2); oc.maximum_f32(A)
        oc.sum_dims(H, 1); oc.sum_dims(H, 2); oc.maximum_bf16(H)
        oc.count_bool(Bm)

    steps = max(1, min(args.steps, 5))
    for _ in range(min(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    gbs = sample_bytes * steps / dt / 1e9
    sample = f"{M}x{cols} column block of each matrix (1/4 of a step), {steps} passes, single thread"
    line = {
        "impl": "reference", "metric": "HBM GB/s, sum/maximum/count on 16384x16384 F32+BF16+Bool (configs[1])",
        "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": dt / steps * 1e3 * (STEP_BYTES / sample_bytes), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: sum(A;dims=1), sum(A;dims=2), maximum(A) on 16384x16384 Float32 and "
                               "BFloat16, count on 16384x16384 Bool; CPU sample scaled to the full step"},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count(),
                         "note": "Base Julia's sum/maximum/count are single-threaded; C restatement oracle/oracle_cpu.c"},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)


# ============================================================================ product arm
def run_product(args, rank, world, local_rank):
    import torch
    import b200arrays as b
    from b200arrays import dtypes as dt

    assert torch.cuda.is_available(), "bench.py needs a B200"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib = b._lib.load()            # fails loudly when the extension is missing
    assert lib.b200_check_device() == 0, "not an sm_100 device"
    dist = None
    # B200_BENCH_FORCE_SHARDED=1: run the N > 1 code path (process group, sharded extras) on ONE GPU — a smoke test
    force_sharded = world == 1 and os.environ.get("B200_BENCH_FORCE_SHARDED") == "1"
    if world > 1 or force_sharded:
        import torch.distributed as dist_
        dist = dist_
        if force_sharded:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    peak, peak_src = measured_peak()
    g = torch.Generator(device=dev); g.manual_seed(0xB200 + 1 + rank)
    # ---- synthetic inputs of configs[1]: F32 uniform [-1,1), BF16 = rounded same, Bool p=0.5
    A32 = torch.rand(M * M, device=dev, generator=g, dtype=torch.float32).mul_(2).sub_(1)
    A16 = A32.to(torch.bfloat16)
    AB = torch.rand(M * M, device=dev, generator=g) < 0.5
    cA32 = b.CuArray.from_torch(A32, (M, M)); cA16 = b.CuArray.from_torch(A16, (M, M)); cAB = b.CuArray.from_torch(AB, (M, M))
    r1_32 = b.CuArray.undef(dt.F32, (1, M)); r2_32 = b.CuArray.undef(dt.F32, (M, 1)); rm_32 = b.CuArray.undef(dt.F32, (1, 1))
    r1_16 = b.CuArray.undef(dt.BF16, (1, M)); r2_16 = b.CuArray.undef(dt.BF16, (M, 1)); rm_16 = b.CuArray.undef(dt.BF16, (1, 1))
    rc = b.CuArray.undef(dt.I64, (1, 1))
    ninf = -np.inf

    def exchange():
        if dist is None:
            return
        # rows are split across ranks (column blocks): dims=2 partial row sums and the scalars need an allreduce.
        # Measured at 2 GPUs: these five small allreduces 0.79 ms per step; packing them into two buffers (Float64 sums,
        # Float32 maxima) costs ten tiny copy kernels and was SLOWER (0.86 ms); torch's _coalescing_manager raises on
        # mixed dtypes and left the process group in a state that corrupted later collectives.  So: five plain calls.
        dist.all_reduce(r2_32.typed()); dist.all_reduce(r2_16.typed())
        dist.all_reduce(rm_32.typed(), op=dist.ReduceOp.MAX); dist.all_reduce(rm_16.typed(), op=dist.ReduceOp.MAX)
        dist.all_reduce(rc.typed())

    def step():
        # interleave the matrices so no input is still L2-resident when it is read again
        b.mapreducedim_(b.identity, b.add_sum, r1_32, cA32, init=0)
        b.mapreducedim_(b.identity, b.add_sum, r1_16, cA16, init=0)
        b.mapreducedim_(b.identity, b.add_sum, rc, cAB, init=0)          # count
        b.mapreducedim_(b.identity, b.add_sum, r2_32, cA32, init=0)
        b.mapreducedim_(b.identity, b.add_sum, r2_16, cA16, init=0)
        b.mapreducedim_(b.identity, b.max_, rm_32, cA32, init=ninf)
        b.mapreducedim_(b.identity, b.max_, rm_16, cA16, init=ninf)
        exchange()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.b200_reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = int(lib.b200_launch_count())
    if dist is not None:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = STEP_BYTES * world / (ms_per_step * 1e-3) / 1e9

    # ---- per-kernel roofline of the dominant kernel: time each reduction alone (events on the launch stream)
    #      L2 is flushed by writing 256 MB and then READING another 256 MB, so the timed kernel starts with a cold
    #      L2 holding clean lines only (the write-backs of the flush's dirty lines are not billed to it)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_clean = torch.zeros(64 << 20, dtype=torch.int32, device=dev)

    def time_one(fn, reps=10):
        ts = []
        for _ in range(reps):
            flush.zero_()
            flush_clean.max()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); e.synchronize()
            ts.append(s.elapsed_time(e))
        return float(np.mean(ts)), float(min(ts))

    per_op = {}
    if rank == 0:
        ops = {
            "sum(A;dims=1) F32 [reduce_contig_block_kernel]": (lambda: b.mapreducedim_(b.identity, b.add_sum, r1_32, cA32, init=0), F32_BYTES),
            "sum(A;dims=2) F32 [reduce_strided_kernel]": (lambda: b.mapreducedim_(b.identity, b.add_sum, r2_32, cA32, init=0), F32_BYTES),
            "maximum(A) F32 [reduce_contig_block_kernel]": (lambda: b.mapreducedim_(b.identity, b.max_, rm_32, cA32, init=ninf), F32_BYTES),
            "sum(A;dims=1) BF16 [reduce_contig_block_kernel]": (lambda: b.mapreducedim_(b.identity, b.add_sum, r1_16, cA16, init=0), BF16_BYTES),
            "sum(A;dims=2) BF16 [reduce_strided_kernel]": (lambda: b.mapreducedim_(b.identity, b.add_sum, r2_16, cA16, init=0), BF16_BYTES),
            "maximum(A) BF16 [reduce_contig_block_kernel]": (lambda: b.mapreducedim_(b.identity, b.max_, rm_16, cA16, init=ninf), BF16_BYTES),
            "count(A) Bool [reduce_contig_block_kernel]": (lambda: b.mapreducedim_(b.identity, b.add_sum, rc, cAB, init=0), BOOL_BYTES),
        }
        for name, (fn, nbytes) in ops.items():
            mean_ms, min_ms = time_one(fn)
            per_op[name] = {"ms": mean_ms, "min_ms": min_ms, "GB/s": nbytes / mean_ms / 1e6, "frac_of_measured_peak": nbytes / mean_ms / 1e6 / peak}
    del flush, flush_clean

    # ---- e2e: the same step through the public API from pinned HOST buffers
    hA32 = A32.cpu().pin_memory(); hA16 = A16.cpu().pin_memory(); hAB = AB.cpu().pin_memory()
    h2d = hA32.numel() * 4 + hA16.numel() * 2 + hAB.numel()
    d2h = (M * 4 + M * 4 + 4) + (M * 2 + M * 2 + 2) + 8

    def e2e_step():
        dA32 = b.CuArray.from_torch(hA32.to(dev, non_blocking=True), (M, M))
        dA16 = b.CuArray.from_torch(hA16.to(dev, non_blocking=True), (M, M))
        dAB = b.CuArray.from_torch(hAB.to(dev, non_blocking=True), (M, M))
        outs = [b.sum(dA32, dims=1), b.sum(dA32, dims=2), b.sum(dA16, dims=1), b.sum(dA16, dims=2)]
        scal = [b.maximum(dA32), b.maximum(dA16), b.count(dAB)]        # dims=: reads the scalar back (D2H + sync)
        return [o.to_host() for o in outs], scal

    e2e_steps = max(1, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = STEP_BYTES * world / (e2e_ms * 1e-3) / 1e9
    del hA32, hA16, hAB

    clocks = sampler.stop() if rank == 0 else {}
    del A32, A16, AB, cA32, cA16, cAB
    torch.cuda.empty_cache()

    extras = {}
    if not args.no_extras and world == 1 and not force_sharded:
        extras = run_extras(b, dt, torch, dev, peak)
    run_sharded = not args.no_extras and (world > 1 or force_sharded)

    if rank != 0 and not run_sharded:
        return
    # ---- cpu baseline (bounded sample of the same workload, host cores of this box)
    cpu = cpu_baseline_sample() if rank == 0 else None
    dom = max(per_op.items(), key=lambda kv: kv[1]["ms"]) if per_op else (None, None)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
            traffic = tj.get("reduce_strided_kernel_f32_dims2_bytes_per_launch" if (dom[0] and "strided" in dom[0])
                             else "reduce_contig_block_kernel_f32_dims1_bytes_per_launch")
    except Exception:
        pass
    roofline = None
    if dom[0]:
        nbytes = F32_BYTES if "F32" in dom[0] else (BF16_BYTES if "BF16" in dom[0] else BOOL_BYTES)
        ach = nbytes / dom[1]["ms"] / 1e6
        roofline = {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "frac_of_nominal_8TBps": ach / NOMINAL_GBS, "peak_source": peak_src, "traffic": traffic,
                    "algorithmic_bytes_per_launch": nbytes, "per_kernel": per_op}
    line = {
        "metric": "HBM GB/s, sum/maximum/count on 16384x16384 F32+BF16+Bool (configs[1])",
        "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: sum(A;dims=1), sum(A;dims=2), maximum(A) on 16384x16384 Float32 and "
                               "BFloat16, count on 16384x16384 Bool; 7 reductions = 5.10 GB algorithmic bytes per step",
                   "l2": "inputs (256 MB - 1 GB each) exceed the 126 MB L2 and are interleaved between uses; the per-kernel "
                         "roofline timings flush L2 between launches (write 256 MB, then read 256 MB so no dirty lines remain)",
                   "sharding": "column blocks, one 16384x16384 block per rank" if world > 1 else "single GPU",
                   "frac_of_measured_peak": value / world / peak, "frac_of_nominal_8TBps": value / world / NOMINAL_GBS},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "steps": e2e_steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "extras": extras,
    }
    if run_sharded:
        # SURVEY section 8e paths (sharded scan / sample sort / 2^30-per-GPU BF16 reductions) in the same launch, so that the
        # driver's 2/4/8-GPU runs carry them.  They come AFTER the headline measurement and can never cost the JSON line.
        finish_with_sharded_extras(line, rank, lambda: run_extras_sharded(b, dt, torch, dist, dev, rank, world, peak))
    print(json.dumps(line), file=_REAL_STDOUT, flush=True)
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline_sample():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import oracle_cpu as oc
        cols = 2048
        rng = np.random.default_rng(0xB200 + 1)
        A = np.asfortranarray(rng.uniform(-1, 1, size=(M, cols)).astype(np.float32))
        H = np.asfortranarray((A.view(np.uint32) >> 16).astype(np.uint16))
        Bm = np.asfortranarray(rng.random((M, cols)) < 0.5)
        nb = 3 * A.nbytes + 3 * H.nbytes + Bm.nbytes
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            oc.sum_dims(A, 1); oc.sum_dims(A, 2); oc.maximum_f32(A)
            oc.sum_dims(H, 1); oc.sum_dims(H, 2); oc.maximum_bf16(H)
            oc.count_bool(Bm)
        dt_ = time.perf_counter() - t0
        return {"value": nb * reps / dt_ / 1e9, "unit": "GB/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
                "sample": f"{M}x{cols} column block of each matrix (1/8 of a step), {reps} passes, single thread "
                          "(Base Julia's reductions are single-threaded); oracle/oracle_cpu.c"}
    except Exception as e:  # the baseline is a reported number, never the product path
        return {"value": None, "unit": "GB/s", "cores": 1, "kind": "port", "sample": f"unavailable: {e}"}


def run_comparators(b, dt, torch, dev):
    """BASELINE.md §2 baselines timed in the same run on the same inputs' shapes:
    R = the reference's ALGORITHMS restated in CUDA C++ (bench/baselines/ref_algos.cu; the reference is Julia and
    cannot run here), L = library state of practice (torch.sum / cumsum / sort, i.e. CUB)."""
    out = {}
    try:
        sys.path.insert(0, os.path.join(ROOT, "bench", "baselines"))
        import ref_algos as R
        R.lib()
    except Exception as e:
        R = None
        out["reference_algorithms"] = f"unavailable: {e}"
    stream = torch.cuda.current_stream(dev).cuda_stream
    try:
        import cub_baseline as C
        C.lib()
    except Exception as e:
        C = None
        out["cub"] = f"unavailable: {e}"

    def tmed(fn, reps, pre=None):
        ts = []
        for _ in range(reps):
            if pre is not None:
                pre()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); e.synchronize()
            ts.append(s.elapsed_time(e))
        return float(np.median(ts))

    g = torch.Generator(device=dev); g.manual_seed(0xB200 + 7)
    # configs[1] F32 matrix
    A = torch.rand(M * M, device=dev, generator=g, dtype=torch.float32)
    At = A.view(M, M)                      # torch is row-major: dim 1 of the view is the contiguous (Julia dim 1) axis
    r = torch.empty(M, device=dev, dtype=torch.float32)
    part = torch.empty(M * 4096, device=dev, dtype=torch.float32)
    for name, tdim, pre, red, post in (("sum(A;dims=1) 16384^2 F32", 1, 1, M, M), ("sum(A;dims=2) 16384^2 F32", 0, M, M, 1)):
        d = {}
        fn = lambda: torch.sum(At, dim=tdim, out=r)
        fn(); d["torch_ms"] = tmed(fn, 5)
        if R is not None:
            fn = lambda: R.sum_f32(A.data_ptr(), r.data_ptr(), part.data_ptr(), pre, red, post, stream)
            fn(); d["reference_algorithm_ms"] = tmed(fn, 5)
        out[name] = d
    del A, At, r
    n = 1 << 30
    V = torch.rand(n, device=dev, generator=g, dtype=torch.float32)
    r1 = torch.empty(1, device=dev, dtype=torch.float32)
    d = {}
    fn = lambda: torch.sum(V); fn(); d["torch_ms"] = tmed(fn, 5)
    if C is not None:
        fn = C.call("cub_sum_f32", torch, dev, V.data_ptr(), r1.data_ptr(), n=n, stream=stream); fn(); d["cub_ms"] = tmed(fn, 5)
    if R is not None:
        fn = lambda: R.sum_f32(V.data_ptr(), r1.data_ptr(), part.data_ptr(), 1, n, 1, stream); fn(); d["reference_algorithm_ms"] = tmed(fn, 5)
    out["sum 2^30 Float32"] = d
    O = torch.empty_like(V)
    d = {}
    fn = lambda: torch.cumsum(V, 0, out=O); fn(); d["torch_ms"] = tmed(fn, 3)
    if C is not None:
        fn = C.call("cub_inclusive_sum_f32", torch, dev, V.data_ptr(), O.data_ptr(), n=n, stream=stream); fn(); d["cub_ms"] = tmed(fn, 3)
    if R is not None:
        scratch = torch.empty(n // 512 + 4096, device=dev, dtype=torch.float32)
        fn = lambda: R.cumsum(V.data_ptr(), O.data_ptr(), scratch.data_ptr(), n, stream); fn(); d["reference_algorithm_ms"] = tmed(fn, 3)
        del scratch
    out["cumsum 2^30 Float32"] = d
    del O
    d = {}
    fn = lambda: torch.sort(V); fn(); d["torch_ms"] = tmed(fn, 2)
    d["torch_note"] = "torch.sort returns values AND Int64 indices (a pairs sort)"
    if C is not None:
        O = torch.empty_like(V)
        fn = C.call("cub_sort_keys_f32", torch, dev, V.data_ptr(), O.data_ptr(), n=n, stream=stream); fn(); d["cub_sort_keys_ms"] = tmed(fn, 3)
        vi = torch.arange(n, device=dev, dtype=torch.int32); vo = torch.empty_like(vi)
        fn = C.call("cub_sort_pairs_f32_u32", torch, dev, V.data_ptr(), O.data_ptr(), vi.data_ptr(), vo.data_ptr(), n=n, stream=stream)
        fn(); d["cub_sort_pairs_u32_payload_ms"] = tmed(fn, 3)
        del O, vi, vo
    out["sort 2^30 Float32"] = d
    del V, part
    torch.cuda.empty_cache()
    if C is not None:
        U = torch.randint(-2**31, 2**31 - 1, (n,), device=dev, generator=g, dtype=torch.int32)
        O = torch.empty_like(U)
        fn = C.call("cub_sort_keys_u32", torch, dev, U.data_ptr(), O.data_ptr(), n=n, stream=stream); fn()
        out["sort! 2^30 UInt32 (CUB DeviceRadixSort::SortKeys)"] = {"cub_ms": tmed(fn, 3), "Gkeys/s": n / tmed(fn, 3) / 1e6}
        del U, O
        torch.cuda.empty_cache()
    if R is not None:
        U = torch.randint(-2**31, 2**31 - 1, (n,), device=dev, generator=g, dtype=torch.int32)
        K = U.clone()
        fn = lambda: R.bitonic_sort_u32(K.data_ptr(), n, stream)
        fn()
        ok = bool((K.view(torch.uint32)[1:1 << 20].to(torch.int64) >= K.view(torch.uint32)[:(1 << 20) - 1].to(torch.int64)).all().item())
        out["sort! 2^30 UInt32"] = {"reference_algorithm_ms": tmed(fn, 1, pre=lambda: K.copy_(U)), "reference_algorithm_sorted_prefix_ok": ok,
                                    "launches": 240}
        del U, K
    torch.cuda.empty_cache()
    out["_labels"] = {"reference_algorithm": "restatement of CUDA.jl's algorithm + launch structure in CUDA C++ (nvcc, not Julia-JIT): "
                                            "bench/baselines/ref_algos.cu", "torch": "torch 2.11 (CUB) library comparator",
                      "cub": "CUB DeviceReduce / DeviceScan / DeviceRadixSort of the CUDA 12.9 toolkit called directly "
                             "(bench/baselines/cub_baseline.cu)"}
    return out


def finish_with_sharded_extras(line, rank, fn, timeout=None):
    """Runs fn() (the sharded extras, collective: every rank calls it), stores its result in line["extras"], prints the
    JSON line on rank 0 and ends the process with status 0 — whatever fn does: an exception is recorded in the line, and a
    watchdog prints the line without the extras if fn has not returned after `timeout` seconds (a rank stuck in a
    collective must not hold the headline numbers back).  Never returns."""
    timeout = float(os.environ.get("B200_BENCH_SHARDED_TIMEOUT", "240")) if timeout is None else timeout
    lock = threading.Lock()
    state = {"printed": False}

    def emit(extras):
        with lock:
            if state["printed"]:
                return
            state["printed"] = True
            if rank == 0:
                line["extras"] = {"sharded": extras}
                print(json.dumps(line), file=_REAL_STDOUT, flush=True)
        os._exit(0)       # not sys.exit: a failed extra may have left a collective half done

    dog = threading.Timer(timeout, emit, args=("timeout: the headline numbers above are complete",))
    dog.daemon = True
    dog.start()
    try:
        res = fn()
    except BaseException as e:  # reported, never fatal for the headline
        res = f"failed: {type(e).__name__}: {e}"
    emit(res)


def run_extras_sharded(b, dt, torch, dist, dev, rank, world, peak):
    """configs[2..4] on `world` GPUs (one process per GPU, NCCL): cumsum and sort! / sortperm of 2^30 elements in total
    (strong scaling), mapreduce(abs2, +) / extrema over 2^30 BFloat16 per GPU (weak scaling, 2^33 on 8 GPUs).  Device
    time between CUDA events on every rank, max over ranks.  Every strong-scaling item carries `ms_1gpu` — the SAME
    total problem on one GPU through the single-GPU entry points, timed on rank 0 in this very run — and
    `speedup_vs_1gpu`; every result is verified on the device (`verified`)."""
    from b200arrays import multigpu as mg
    out = {}

    def timed(fn, reps=3):
        fn()
        best = None
        for _ in range(reps):
            dist.barrier(); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); e.synchronize()
            t = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t.item()) if best is None else min(best, float(t.item()))
        return best

    def one_gpu(make, fn, reps=2, restore=False):
        """rank 0 times fn(x) on x = make() (the whole problem on one GPU); the others wait.  restore: fn works in place
        (sort!), so x is reset from a pristine copy before every run, outside the timed region."""
        ms = None
        if rank == 0:
            try:
                x = make()
                keep = x.copy() if restore else None
                fn(x)
                ts = []
                for _ in range(reps):
                    if restore:
                        x.buf.copy_(keep.buf)
                    torch.cuda.synchronize()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(); fn(x); e.record(); e.synchronize()
                    ts.append(s.elapsed_time(e))
                ms = min(ts)
                del x, keep
            except Exception as ex:
                ms = None
                print(f"1-GPU baseline failed: {ex}", file=sys.stderr)
            torch.cuda.empty_cache()
        dist.barrier()
        return ms

    def all_ok(flag):
        t = torch.tensor([1 if flag else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    def sorted_across_ranks(res, n_total, et_torch, rev=False):
        """sorted inside the rank, boundary keys ordered across ranks, slice lengths sum to N (all on the device)."""
        t = res.typed()
        if et_torch == torch.uint32:
            t = t.view(torch.int32) ^ -2**31                      # order-preserving signed view
        ok = bool((t[1:] >= t[:-1]).all()) if t.numel() > 1 else True
        edge = torch.zeros(3, device=dev, dtype=torch.float64)     # first, last, count
        if t.numel():
            edge[0] = t[0].double(); edge[1] = t[-1].double()
        edge[2] = t.numel()
        edges = [torch.empty_like(edge) for _ in range(world)]
        dist.all_gather(edges, edge)
        last = None
        total = 0
        for e in edges:
            total += int(e[2].item())
            if int(e[2].item()) == 0:
                continue
            if last is not None and float(e[0].item()) < last:
                ok = False
            last = float(e[1].item())
        return all_ok(ok and total == n_total)

    def item(name, fn, fmt, reps=3, base=None, verify=None):
        # every rank runs the same code on the same shapes, so a failure is the same on all of them: record it and go on
        try:
            ms = timed(fn, reps)
            d = fmt(ms)
            if verify is not None:
                d["verified"] = verify()
            if base is not None:
                ms1 = one_gpu(*base)
                if rank == 0 and ms1:
                    d["ms_1gpu"] = ms1
                    d["speedup_vs_1gpu"] = ms1 / ms
            out[name] = d
        except Exception as e:
            import traceback
            out[name] = {"failed": f"{type(e).__name__}: {e}", "where": traceback.format_exc().strip().splitlines()[-6:]}

    lg = int(os.environ.get("B200_BENCH_SHARDED_LOG2N", "30"))      # smoke tests use a smaller size; the names below say 2^30
    total = 1 << lg
    n = total // world
    g = torch.Generator(device=dev); g.manual_seed(0xB200 + 100 + rank)
    g1 = torch.Generator(device=dev); g1.manual_seed(0xB200 + 99)

    V = torch.rand(n, device=dev, generator=g, dtype=torch.float32)
    cV = b.CuArray.from_torch(V)
    holder = {}

    def cumsum_f32():
        holder["r"] = mg.dist_accumulate(b.add, cV)

    def verify_cumsum(shard_t, acc_dtype, tol):
        """last element of every rank's output == sum of all elements up to the end of that rank (computed independently
        with torch in Float64 / Int64); exact for integers, relative tolerance for floats."""
        r = holder["r"].typed()
        part = shard_t.to(acc_dtype).sum()
        parts = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(parts, part)
        want = sum(float(p.item()) if tol else int(p.item()) for p in parts[:rank + 1])
        got = float(r[-1].item()) if tol else int(r[-1].item())
        ok = abs(got - want) <= tol * abs(want) if tol else got == want
        if not ok:
            print(f"[rank {rank}] cumsum check: got {got!r} want {want!r} parts {[p.item() for p in parts]}", file=sys.stderr, flush=True)
        return all_ok(ok)

    def mk_f32():
        return b.CuArray.from_torch(torch.rand(total, device=dev, generator=g1, dtype=torch.float32))

    item("cumsum 2^30 Float32 total", cumsum_f32,
         lambda ms: {"ms": ms, "GB/s": 8 * total / ms / 1e6, "per_gpu_frac_of_measured_peak": 8 * total / ms / 1e6 / world / peak,
                     "note": "shard-local reduce + all-gather of the shard totals + scan seeded on the device with the totals "
                             "before this rank (b200_scan_carry): 3N traffic for 2N algorithmic bytes, no host sync"},
         base=(mk_f32, lambda x: b.accumulate_(b.add, x, x)),
         verify=lambda: verify_cumsum(V, torch.float64, 4 * 30 * 2.0 ** -23))

    def sort_f32():
        holder["r"] = mg.dist_sort(cV)[0]

    item("sort! 2^30 Float32 total", sort_f32, lambda ms: {"ms": ms, "Gkeys/s": total / ms / 1e6},
         base=(mk_f32, lambda x: b.sort_(x), 2, True), verify=lambda: sorted_across_ranks(holder["r"], total, torch.float32))
    item("sortperm 2^30 Float32 total", lambda: mg.dist_sortperm(cV), lambda ms: {"ms": ms, "Gkeys/s": total / ms / 1e6}, reps=2,
         base=(mk_f32, lambda x: b.sortperm(x)))
    holder.clear()
    del V, cV
    torch.cuda.empty_cache()
    I = torch.randint(-2**20, 2**20, (n,), device=dev, generator=g, dtype=torch.int64)
    cI = b.CuArray.from_torch(I)

    def cumsum_i64():
        holder["r"] = mg.dist_accumulate(b.add, cI)

    item("cumsum 2^30 Int64 total", cumsum_i64,
         lambda ms: {"ms": ms, "GB/s": 16 * total / ms / 1e6, "per_gpu_frac_of_measured_peak": 16 * total / ms / 1e6 / world / peak},
         base=(lambda: b.CuArray.from_torch(torch.randint(-2**20, 2**20, (total,), device=dev, generator=g1, dtype=torch.int64)),
               lambda x: b.accumulate_(b.add, x, x)),
         verify=lambda: verify_cumsum(I, torch.int64, 0))
    holder.clear()
    del I, cI
    torch.cuda.empty_cache()
    U = torch.randint(-2**31, 2**31 - 1, (n,), device=dev, generator=g, dtype=torch.int32)
    cU = b.CuArray.from_torch(U, (n,), dt.U32)

    def sort_u32():
        holder["r"] = mg.dist_sort(cU)[0]

    item("sort! 2^30 UInt32 total", sort_u32, lambda ms: {"ms": ms, "Gkeys/s": total / ms / 1e6},
         base=(lambda: b.CuArray.from_torch(torch.randint(-2**31, 2**31 - 1, (total,), device=dev, generator=g1, dtype=torch.int32), (total,), dt.U32),
               lambda x: b.sort_(x), 2, True),
         verify=lambda: sorted_across_ranks(holder["r"], total, torch.uint32))
    holder.clear()
    del U, cU
    torch.cuda.empty_cache()
    H = torch.randn(total, device=dev, generator=g, dtype=torch.float32).to(torch.bfloat16)
    cH = b.CuArray.from_torch(H)
    item(f"mapreduce(abs2,+) 2^30 BFloat16 per GPU x {world}", lambda: mg.dist_mapreduce(b.abs2, b.add, cH),
         lambda ms: {"ms": ms, "GB/s": 2.0 * total * world / ms / 1e6, "per_gpu_frac_of_measured_peak": 2.0 * total / ms / 1e6 / peak})
    item(f"extrema 2^30 BFloat16 per GPU x {world}", lambda: mg.dist_extrema(cH),
         lambda ms: {"ms": ms, "GB/s": 2.0 * total * world / ms / 1e6, "per_gpu_frac_of_measured_peak": 2.0 * total / ms / 1e6 / peak,
                     "note": "one fused (min, max) pass per shard, one all-gather of the G pairs"})
    return out


def run_extras(b, dt, torch, dev, peak):
    """North-star configurations beyond configs[1], measured in the same run (1 GPU)."""
    out = {}

    def tmin(fn, reps, pre=None):
        ts = []
        for _ in range(reps):
            if pre is not None:
                pre()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); e.synchronize()
            ts.append(s.elapsed_time(e))
        return float(np.median(ts)), float(min(ts))

    n = 1 << 30
    g = torch.Generator(device=dev); g.manual_seed(0xB200 + 3)
    V = torch.rand(n, device=dev, generator=g, dtype=torch.float32)
    cV = b.CuArray.from_torch(V)
    R = b.CuArray.undef(dt.F32, (1,))
    fn = lambda: b.mapreducedim_(b.identity, b.add_sum, R, cV, init=0)
    fn(); med, mn = tmin(fn, 10)
    out["sum 2^30 Float32"] = {"ms": med, "GB/s": 4 * n / med / 1e6, "frac_of_measured_peak": 4 * n / med / 1e6 / peak,
                               "frac_of_nominal_8TBps": 4 * n / med / 1e6 / NOMINAL_GBS}
    O = b.CuArray.undef(dt.F32, (n,))
    fn = lambda: b.accumulate_(b.add, O, cV)
    fn(); med, mn = tmin(fn, 5)
    out["cumsum 2^30 Float32"] = {"ms": med, "GB/s": 8 * n / med / 1e6, "frac_of_measured_peak": 8 * n / med / 1e6 / peak,
                                  "frac_of_nominal_8TBps": 8 * n / med / 1e6 / NOMINAL_GBS}
    del O
    # sort! / sortperm Float32 (keys restored from a pristine copy outside the timed region)
    K = cV.copy()
    restore = lambda: K.buf.copy_(cV.buf)
    fn = lambda: b.sort_(K)
    restore(); fn(); med, mn = tmin(fn, 3, pre=restore)
    out["sort! 2^30 Float32"] = {"ms": med, "Gkeys/s": n / med / 1e6, "algorithmic_GB/s": 2 * 4 * n * 4 / med / 1e6,
                                 "frac_of_radix_roofline_8TBps": 2 * 4 * n * 4 / med / 1e6 / NOMINAL_GBS,
                                 "frac_of_radix_roofline_measured_peak": 2 * 4 * n * 4 / med / 1e6 / peak}
    del K
    fn = lambda: b.sortperm(cV)
    fn(); med, mn = tmin(fn, 3)
    out["sortperm 2^30 Float32"] = {"ms": med, "Gkeys/s": n / med / 1e6}
    fn = lambda: b.partialsort(cV, n // 2)          # radix select (sampled bracket + one-read compaction), input untouched
    fn(); med, mn = tmin(fn, 5)
    out["partialsort(c, n/2) 2^30 Float32"] = {"ms": med, "GB/s_counting_one_read_of_the_input": 4 * n / med / 1e6}
    del V, cV
    torch.cuda.empty_cache()
    U = torch.randint(-2**31, 2**31 - 1, (n,), device=dev, generator=g, dtype=torch.int32)
    cU = b.CuArray.from_torch(U, (n,), dt.U32)
    K = cU.copy()
    restore = lambda: K.buf.copy_(cU.buf)
    fn = lambda: b.sort_(K)
    restore(); fn(); med, mn = tmin(fn, 3, pre=restore)
    out["sort! 2^30 UInt32"] = {"ms": med, "Gkeys/s": n / med / 1e6, "algorithmic_GB/s": 2 * 4 * n * 4 / med / 1e6,
                                "frac_of_radix_roofline_8TBps": 2 * 4 * n * 4 / med / 1e6 / NOMINAL_GBS,
                                "frac_of_radix_roofline_measured_peak": 2 * 4 * n * 4 / med / 1e6 / peak}
    del K, U, cU
    torch.cuda.empty_cache()
    I = torch.randint(-2**20, 2**20, (n,), device=dev, generator=g, dtype=torch.int64)
    cI = b.CuArray.from_torch(I)
    O = b.CuArray.undef(dt.I64, (n,))
    fn = lambda: b.accumulate_(b.add, O, cI)
    fn(); med, mn = tmin(fn, 5)
    out["cumsum 2^30 Int64"] = {"ms": med, "GB/s": 16 * n / med / 1e6, "frac_of_measured_peak": 16 * n / med / 1e6 / peak,
                                "frac_of_nominal_8TBps": 16 * n / med / 1e6 / NOMINAL_GBS}
    del I, cI, O
    torch.cuda.empty_cache()
    try:
        out["configs[0] 2^24 Float32: Base Julia CPU path (port) beside the graft"] = run_configs0(b, dt, torch, dev)
    except Exception as e:
        out["configs[0] 2^24 Float32: Base Julia CPU path (port) beside the graft"] = {"error": repr(e)}
    try:
        out["comparators"] = run_comparators(b, dt, torch, dev)
    except Exception as e:  # baselines are reported numbers, never the product path
        out["comparators"] = {"error": repr(e)}
    return out


def run_configs0(b, dt, torch, dev):
    """BASELINE.json configs[0]: sum / cumsum / sort! / sortperm on a 2^24-element Float32 Array via Base Julia's CPU
    path — here its C restatement (oracle/oracle_cpu.c, single thread like Base) — next to the same calls on the GPU
    through the public API, data resident.  For sort both CPU algorithms SURVEY section 8(d) asks for are timed: LSD radix
    (what Julia >= 1.9 picks for large Float32 vectors) and a comparison sort."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_cpu as oc
    n = 1 << 24
    rng = np.random.default_rng(0xB200)
    a = rng.random(n).astype(np.float32)

    def cpu(fn, reps=2):
        best = None
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); t = (time.perf_counter() - t0) * 1e3
            best = t if best is None else min(best, t)
        return best

    def gpu(fn, reps=5, pre=None):
        ts = []
        for _ in range(reps + 1):
            if pre is not None:
                pre()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); e.synchronize()
            ts.append(s.elapsed_time(e))
        return float(np.median(ts[1:]))

    cA = b.CuArray(a)
    R = b.CuArray.undef(dt.F32, (1,)); O = b.CuArray.undef(dt.F32, (n,)); K = cA.copy()
    res = {
        "sum": {"cpu_ms": cpu(lambda: oc.sum_f32(a)), "gpu_ms": gpu(lambda: b.mapreducedim_(b.identity, b.add_sum, R, cA, init=0))},
        "cumsum": {"cpu_ms": cpu(lambda: oc.cumsum(a)), "gpu_ms": gpu(lambda: b.accumulate_(b.add, O, cA))},
        "sort!": {"cpu_radix_ms": cpu(lambda: oc.sort(a)), "cpu_comparison_sort_ms": cpu(lambda: oc.sort_cmp(a), reps=1),
                  "gpu_ms": gpu(lambda: b.sort_(K), pre=lambda: K.buf.copy_(cA.buf))},
        "sortperm": {"cpu_ms": cpu(lambda: oc.sortperm(a), reps=1), "gpu_ms": gpu(lambda: b.sortperm(cA))},
        "cpu": {"cores": 1, "host_cores": os.cpu_count(), "kind": "port",
                "note": "Base Julia's sum / cumsum / sort! / sortperm are single-threaded (Threads.nthreads() is irrelevant); "
                        "oracle/oracle_cpu.c restates their algorithms; the sort timings include one copy of the input"},
    }
    return res


_REAL_STDOUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the contract is ONE JSON line on stdout: libraries that write to file descriptor 1 themselves (NCCL prints its
    # version banner there when NCCL_DEBUG is set) are sent to stderr; the JSON line goes to the real stdout
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_product(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
