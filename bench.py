#!/usr/bin/env python
"""bench.py -- headline benchmark of the polynomial-ring hot path (BASELINE.json metric
"NTTs/s at N=4096; bootstraps/s; at 1/2/4/8 B200 vs host-CPU reference").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--driver torchrun|capi]

Headline workload (BASELINE.json configs[1]): batched forward negacyclic NTT, N = 4096, 64-bit NTT-friendly prime
q = 1125899906826241 (the reference's own bench prime, primus_ntt/benches/bench_u64.rs:8), batch 65536 polynomials (2 GiB)
per GPU.  A "step" is one pass of the hot path over the whole batch (one kernel launch, in place).

  value        whole-job NTTs/s with the batch resident in HBM (CUDA events, max over ranks)
  e2e          the same metric through the reference-facing C-ABI host-slice call (pfhe_ntt64_transform_slices): pinned HOST
               buffer -> H2D -> kernel -> D2H, all inside the timed region; e2e.pageable = the same with pageable memory
               (what a Rust `&mut [u64]` is)
  roofline     algorithmic bytes (2*N*8 per NTT) / kernel time vs the measured HBM copy peak; frac_sustained = the same over a
               >= 1 s back-to-back run (the board sits at its power cap there)
  cpu_baseline the CPU oracle (restated reference scalar path, OpenMP over the batch) on this box's cores
  bootstrap    the second half of the metric: blind-rotation bootstraps/s on BASELINE config 5 (n = 512, N = 1024, u32 q = 132120577,
               base 2^7; 10,000 ciphertexts split over the ranks), with its own roofline / cpu_baseline / e2e
  extra        secondary workloads of the same path, each parity-checked against the oracle outside the timed region

Multi-GPU: one process per GPU (torchrun), the batch is sharded (independent units, no collective on the data path; NCCL only for
the barrier / max-over-ranks of the timing) -> "scaling": "weak" for the NTT line; the bootstrap batch is fixed at 10,000 in total
("strong").  `--driver capi` reproduces the same measurement from ONE process through pfhe_multi_* (no torch.distributed).

`--impl reference` times the reference arm: the reference's CPU algorithm (oracle port, all host threads) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = 12
N = 1 << LOG_N
Q = 1125899906826241
BATCH = 65536
BYTES_PER_NTT = 2 * N * 8            # read + write, SURVEY.md 8(d)
MODMULS_PER_NTT = (N // 2) * LOG_N   # butterflies
FP64_PER_NTT = 944 * 256             # FP64 instructions of the lazy-fold kernel: 96 butterflies x 8 + 3 folds x 16 x 3 + 32 conversions, per thread
FP64_PEAK = 1.75e13                  # measured DFMA/DMUL/DADD rate of one B200 (profiles/r01_ubench_pipes.log)
METRIC = "NTTs/s at N=4096 (u64 forward negacyclic NTT, batch 65536 per GPU)"
UNIT = "NTT/s"
CONFIG = {"workload": "batched forward NTT, N=4096, q=1125899906826241 (50-bit), batch 65536 polys (2 GiB) per GPU, in place",
          "l2": "inputs (2 GiB) exceed L2 (126 MB): no flush needed between timed iterations",
          "sharding": "independent polynomials split across GPUs, no data-path collective"}
# BASELINE config 5
BR_Q, BR_LOGN, BR_NLWE, BR_LOGB, BR_TOTAL = 132120577, 10, 512, 7, 10000
BR_MODMULS = BR_NLWE * (2 * 3 * 5120 + 4 * 3 * 1024 + 2 * 5632)   # 27.8 M modular multiplications per bootstrap


def host_threads() -> int:
    """Cores this process may use.  torchrun exports OMP_NUM_THREADS=1, which is NOT the size of the box: the reference arm and the
    CPU baseline use every core in the affinity mask (the oracle takes the thread count explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _c3_primes():
    """8 primes just below 2^50 with q = 1 mod 2^15 (config C3)."""
    def is_prime(n):
        d, r = n - 1, 0
        while d % 2 == 0:
            d //= 2; r += 1
        for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            x = pow(a, d, n)
            if x in (1, n - 1):
                continue
            for _ in range(r - 1):
                x = x * x % n
                if x == n - 1:
                    break
            else:
                return False
        return True
    out, c = [], (1 << 50) - (1 << 15) + 1
    while len(out) < 8:
        if is_prime(c):
            out.append(c)
        c -= 1 << 15
    return out


def _gpu_numa_node(gpu_index: int):
    """NUMA node of the GPU (None when the platform does not expose one, e.g. a single-node VM) and a pin of this process to it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = gpu_index
        if vis and all(p.strip().isdigit() for p in vis.split(",")):
            idx = int(vis.split(",")[gpu_index])
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0:
            return {"node": None, "host_nodes": len(nodes), "note": "PCI device reports numa_node = -1 (no affinity exposed by this platform)"}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"node": node, "host_nodes": len(nodes), "pinned_cpus": len(cpus)}
    except Exception as ex:
        return {"node": None, "note": repr(ex)[:120]}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons through NVML (in-process thread, ~2 ms period) while timed regions run."""
    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.stop_flag, self.thread, self.mx = gpu_index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.gpu
            if vis and all(p.strip().isdigit() for p in vis.split(",")):
                idx = int(vis.split(",")[self.gpu])
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self.stop_flag:
                    try:
                        clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        rs = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0
                        self.rows.append((time.perf_counter(), float(clk), int(rs), pw))
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def window(self, t_begin, t_end):
        inside = [r for r in self.rows if t_begin <= r[0] <= t_end]
        note = "sampled inside the timed region"
        if len(inside) < 3:
            inside, note = list(self.rows), "timed region shorter than 3 samples: all samples of this run so far used"
        sm = [r[1] for r in inside]
        reasons = set()
        for r in inside:
            for name, bit in self.REASONS.items():
                if r[2] & bit:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(reasons),
                "power_w_max": max((r[3] for r in inside), default=None), "samples": len(sm), "note": note}

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)


# ---------------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baselines (the ONLY places that touch oracle/)
# ---------------------------------------------------------------------------------------------------------------------------
def cpu_ntt(sample_batch: int, reps: int):
    """Forward NTT on the host cores: the AVX-512 IFMA restatement of the reference's own fast path when the CPU has it (that is what
    U64NttTable dispatches to for q < 2^50, primus_ntt/src/ntt/prime64/table.rs:166-231), else the scalar Harvey port.
    Returns (NTT/s, threads, seconds, kind note, scalar NTT/s)."""
    import numpy as np
    from oracle import oracle as O
    t = O.U64NttTable(LOG_N, Q)
    threads = host_threads()
    rng = np.random.default_rng(0x5EED0002)
    x = rng.integers(0, Q, (sample_batch, N), dtype=np.uint64)

    def best_of(fn):
        fn(x[:min(256, sample_batch)].copy(), threads)  # warm
        best = 1e30
        for _ in range(reps):
            y = x.copy()
            t0 = time.perf_counter()
            fn(y, threads)
            best = min(best, time.perf_counter() - t0)
        return best
    scalar = best_of(t.forward_batch)
    if t.simd_supported():
        simd = best_of(t.forward_batch_simd)
        return sample_batch / simd, threads, simd, CPU_KIND_SIMD, sample_batch / scalar
    return sample_batch / scalar, threads, scalar, CPU_KIND_SCALAR, sample_batch / scalar


CPU_KIND_SIMD = ("AVX-512 IFMA restatement (oracle/pfhe_oracle_avx512.c) of the reference's HEXL-style back-end "
                 "(primus_ntt/src/ntt/prime64/avx512/{transform,stages,butterfly}.rs, BIT_SHIFT = 52), the path the reference itself selects on this CPU for q < 2^50; "
                 "bit-identical to the scalar port (tests/test_oracle.py); the Rust reference cannot be built here (no cargo)")
CPU_KIND_SCALAR = ("C restatement of the reference's SCALAR Harvey path (this CPU has no AVX-512 IFMA, so the reference would not use its HEXL-style back-end either; "
                   "the Rust reference cannot be built here: no cargo)")


def cpu_bootstrap(sample: int):
    import numpy as np
    from oracle import oracle as O
    threads = host_threads()
    ot = O.U32NttTable(BR_LOGN, BR_Q)
    ob = O.ApproxSignedBasis(BR_Q, BR_LOGB, None, 32)
    lv = ob.decompose_length()
    rng = np.random.default_rng(0x5EED0005)
    n = 1 << BR_LOGN
    bsk = rng.integers(0, BR_Q, BR_NLWE * 2 * lv * 2 * n, dtype=np.uint64).astype(np.uint32)
    lwe = rng.integers(0, 2 * n, (sample, BR_NLWE + 1), dtype=np.uint64).astype(np.uint32)
    tv = rng.integers(0, BR_Q, n, dtype=np.uint64).astype(np.uint32)
    t0 = time.perf_counter()
    O.blind_rotate(ot, ob, bsk, BR_NLWE, lwe, tv, batch=sample, threads=threads)
    dt = time.perf_counter() - t0
    return sample / dt, threads, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = 16384
    times = []
    import numpy as np
    from oracle import oracle as O
    t = O.U64NttTable(LOG_N, Q)
    threads = host_threads()
    simd = t.simd_supported()
    fwd = t.forward_batch_simd if simd else t.forward_batch
    rng = np.random.default_rng(0x5EED0002)
    x = rng.integers(0, Q, (sample, N), dtype=np.uint64)
    for i in range(args.warmup + args.steps):
        y = x.copy()
        t0 = time.perf_counter()
        fwd(y, threads)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    y = x.copy(); t0 = time.perf_counter(); t.forward_batch(y, threads); scalar_value = sample / (time.perf_counter() - t0)
    bs_value, _, bs_dt = cpu_bootstrap(2 * threads)
    kind_note = CPU_KIND_SIMD if simd else CPU_KIND_SCALAR
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": dict(CONFIG, reference_sample=f"{sample} polynomials per step (bounded sample of the 65536-poly batch)"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} NTTs/step x {args.steps} steps, OpenMP over the batch, {threads} threads "
                                       f"(sched_getaffinity; OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS', 'unset')} ignored)",
                             "note": kind_note, "scalar_port_value": scalar_value},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "bootstrap": {"metric": "bootstraps/s (blind rotation n=512, N=1024, u32 q=132120577, base 2^7)", "value": bs_value,
                          "unit": "bootstrap/s", "cores": threads, "kind": "port",
                          "sample": f"{2 * threads} ciphertexts x 512 CMux steps, {bs_dt:.2f} s; scalar u32 port (the reference's u32 AVX-512 back-end is not restated)"}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--driver", default="torchrun", choices=["torchrun", "capi"],
                    help="capi: ONE process drives --gpus devices through pfhe_multi_* (no torch.distributed)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary (extra) measurements")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--sustain-s", type=float, default=1.2, help="length of the sustained-roofline run (seconds, 0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.driver == "capi":
        return run_capi_driver(args)

    import numpy as np
    import torch
    import primus_fhe_b200 as P
    from primus_fhe_b200.shard import shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = _gpu_numa_node(local_rank)
    dist = None
    # stdout carries exactly ONE JSON line: anything libraries print at the C level (NCCL's version banner ...) is sent to
    # stderr by swapping the descriptors for the duration of the run; the line itself goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    checks = {}
    batch = args.batch
    table = P.U64NttTable(LOG_N, Q, device=local_rank)
    g = torch.Generator(device="cuda"); g.manual_seed(0x5EED0002 + rank)
    data = torch.randint(0, Q, (batch, N), dtype=torch.int64, device="cuda", generator=g)
    sample_rows = list(range(8)) + list(range(batch - 8, batch))

    def parity_headline(tag):
        """Outside the timed region: one step on the timed buffer, first/last 8 polynomials vs the oracle, bit for bit."""
        from oracle import oracle as O
        before = data[sample_rows].cpu().numpy().view(np.uint64).copy()
        table.forward_batch(data)
        torch.cuda.synchronize()
        after = data[sample_rows].cpu().numpy().view(np.uint64)
        want = before.copy(); O.U64NttTable(LOG_N, Q).forward_batch(want, 1)
        ok = bool(np.array_equal(after, want))
        checks[tag] = ok
        if not ok:
            raise SystemExit(f"bench.py: parity check '{tag}' FAILED - refusing to time a wrong kernel")

    # ---- device-resident throughput (value) --------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    parity_headline("ntt_fwd_n4096_before_timing")
    for _ in range(args.warmup):
        table.forward_batch(data)
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = P.launch_count()
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    e_all0.record()
    for a, b in evs:
        a.record()
        table.forward_batch(data)      # repeated forward transforms of transformed data: same work, still valid inputs (< q)
        b.record()
    e_all1.record()
    barrier()
    t_end = time.perf_counter()
    launches = P.launch_count() - launches0
    total_ms = e_all0.elapsed_time(e_all1)
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms_max = max_over_ranks(total_ms)
    value = world * batch * args.steps / (total_ms_max * 1e-3)
    clocks = sampler.window(t_begin, t_end) if rank == 0 else None
    parity_headline("ntt_fwd_n4096_after_timing")

    # ---- sustained roofline: >= 1 s of back-to-back launches -----------------------------------------------------
    sustained = None
    if args.sustain_s > 0:
        est = sum(kernel_ms) / len(kernel_ms)
        n_sus = max(200, int(args.sustain_s * 1e3 / est))
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts0 = time.perf_counter()
        s0.record()
        for _ in range(n_sus):
            table.forward_batch(data)
        s1.record()
        barrier()
        ts1 = time.perf_counter()
        sus_ms = max_over_ranks(s0.elapsed_time(s1))
        sustained = {"launches": n_sus, "seconds": sus_ms * 1e-3, "ms_per_launch": sus_ms / n_sus,
                     "value": world * batch * n_sus / (sus_ms * 1e-3),
                     "clocks": sampler.window(ts0, ts1) if rank == 0 else None}

    # ---- end to end through the C-ABI host-slice call ---------------------------------------------------------------
    e2e_batch = batch
    host = torch.empty((e2e_batch, N), dtype=torch.int64).pin_memory()
    host.copy_(data.cpu())
    e2e_steps = max(2, min(args.steps, 5))
    table.transform_slices(host)  # warm (streams, mempool)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        table.transform_slices(host)   # H2D + kernel + D2H of every polynomial, synchronous
    torch.cuda.synchronize()
    e2e_value = world * e2e_batch * e2e_steps / max_over_ranks(time.perf_counter() - t0)
    # pageable host memory: what `transform_slice(&mut [T])` hands over (primus_data/src/traits.rs:20)
    pg_batch = e2e_batch // 4
    pageable = np.empty((pg_batch, N), dtype=np.uint64)
    pageable[:] = host[:pg_batch].numpy().view(np.uint64)
    table.transform_slices(pageable)
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        table.transform_slices(pageable)
    e2e_pageable = world * pg_batch * 2 / max_over_ranks(time.perf_counter() - t0)
    # the caller's own buffer page-locked in place once (pfhe_host_register; ffi/primus_cuda `Pinned`): registration cost reported, not timed
    t0 = time.perf_counter()
    reg = P.registered_host_buffer(pageable)
    reg.__enter__()
    reg_seconds = time.perf_counter() - t0
    try:
        table.transform_slices(pageable)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            table.transform_slices(pageable)
        e2e_registered = world * pg_batch * 2 / max_over_ranks(time.perf_counter() - t0)
    finally:
        reg.__exit__(None, None, None)
    del host, pageable

    # ---- bootstraps/s: BASELINE config 5, every rank runs its share of the 10,000 ciphertexts ---------------------------------
    bootstrap = bench_bootstrap(P, torch, np, world, rank, local_rank, barrier, max_over_ranks, shard_range, sampler, checks)

    if rank != 0:
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel ----------------------------------------------------------------------------------
    peak, peak_src = _peaks()
    avg_kernel_ms = sum(kernel_ms) / len(kernel_ms)
    achieved = batch * BYTES_PER_NTT / (avg_kernel_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("ntt_fwd_u64_n4096", {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "kernel": "ntt_tma_kernel<F64LazyField,N=4096> forward (lazy-fold FP64 butterflies, TMA tensor store)",
                "algorithmic_bytes_per_launch": batch * BYTES_PER_NTT, "avg_launch_ms": avg_kernel_ms,
                "best_launch_ms": min(kernel_ms), "frac_best_launch": batch * BYTES_PER_NTT / (min(kernel_ms) * 1e-3) / 1e9 / peak,
                "note": "achieved/frac use the AVERAGE launch over the timed region; peak is the burst copy bandwidth; frac_sustained is the "
                        "same kernel over a >= 1 s back-to-back run (board at its power cap: see sustained.clocks)",
                "deviations_from_north_star": "q < 2^50 runs exact butterflies on the FP64 pipe (6 FP64 instructions per modular product) instead of "
                                              "integer Montgomery/Shoup (IMAD.WIDE 22/clk/SM, mul.hi.u64 6.5/clk/SM measured); exchanges go through "
                                              "swizzled shared memory + __syncwarp instead of warp shuffles (DESIGN.md 3.2)",
                "modmul": {"butterflies_per_s": batch * MODMULS_PER_NTT / (avg_kernel_ms * 1e-3),
                           "fp64_instr_per_ntt": FP64_PER_NTT, "fp64_pipe_peak_instr_per_s": FP64_PEAK,
                           "frac_of_fp64_pipe": batch * FP64_PER_NTT / (avg_kernel_ms * 1e-3) / FP64_PEAK,
                           "note": "secondary (binding) bound: 944 FP64 instructions per thread x 256 threads per NTT against the measured "
                                   "DFMA rate (profiles/r01_ubench_pipes.log)"}}
    if sustained:
        roofline["frac_sustained"] = batch * BYTES_PER_NTT / (sustained["ms_per_launch"] * 1e-3) / 1e9 / peak
        roofline["sustained"] = sustained

    # ---- CPU baseline (bounded sample) --------------------------------------------------------------------------------
    cpu_value, cores, cpu_s, cpu_note, cpu_scalar = cpu_ntt(8192, 3)
    cpu_baseline = {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"8192 NTTs (1/8 of the batch), best of 3, {cpu_s:.3f} s", "note": cpu_note, "scalar_port_value": cpu_scalar}
    bs_cpu, bs_cores, bs_dt = cpu_bootstrap(2 * cores)
    bootstrap["cpu_baseline"] = {"value": bs_cpu, "unit": "bootstrap/s", "cores": bs_cores, "kind": "port",
                                 "sample": f"{2 * cores} ciphertexts x 512 CMux steps on {bs_cores} threads, {bs_dt:.2f} s (oracle blind rotation, scalar u32 port)"}

    # ---- secondary measurements ---------------------------------------------------------------------------------------------
    extra = {}
    if not args.no_extra:
        try:
            bench_extra(P, torch, np, table, data, g, local_rank, peak, extra, checks)
        except Exception as ex:  # secondary numbers must never hide the headline
            extra["error"] = repr(ex)
    pcie = None
    try:
        pcie = json.load(open(os.path.join(ROOT, "profiles", "r02_pcie_concurrent.json")))
    except Exception:
        pass
    e2e_bytes_per_s = e2e_value / world * 2 * N * 8
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": dict(CONFIG, batch_per_gpu=batch, n_gpus=world),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_batch * N * 8, "d2h_bytes_per_step": e2e_batch * N * 8,
                    "api": "pfhe_ntt64_transform_slices (host-slice shim of NttTable::transform_slice, pinned host memory)",
                    "pageable": {"value": e2e_pageable, "unit": UNIT, "batch": pg_batch,
                                 "note": "same call on pageable memory (numpy buffer): what a Rust `&mut [u64]` caller passes"},
                    "registered": {"value": e2e_registered, "unit": UNIT, "batch": pg_batch, "register_seconds": reg_seconds,
                                   "note": "the same numpy buffer page-locked in place once with pfhe_host_register (not inside the timed region)"},
                    "pcie_gbs_per_direction_per_gpu": e2e_bytes_per_s / 2 / 1e9,
                    "pcie_frac": (e2e_bytes_per_s / 2 / 1e9) / pcie["per_direction_gbs"][str(world)] if pcie and str(world) in pcie.get("per_direction_gbs", {}) else None,
                    "pcie_note": "host<->device bytes (4 GiB per step per GPU) bound this number: see profiles/r02_pcie_concurrent.json for the "
                                 "concurrent per-direction copy bandwidth measured at 1/2/4/8 ranks on the same box type",
                    "numa": numa, "steps": e2e_steps},
            "bootstrap": bootstrap, "parity_checks": checks,
            "gpu_launches": int(launches), "clocks": clocks, "extra": extra}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


def modmul_peak(P, device):
    """u32 Shoup products per second at integer-pipe peak, measured live (pfhe_modmul_microbench kind 2)."""
    blocks, iters = 148 * 8, 4096
    ms = min(P.modmul_microbench(2, blocks, iters, device) for _ in range(3))
    return blocks * 256 * 8 * iters / (ms * 1e-3)


def bench_bootstrap(P, torch, np, world, rank, local_rank, barrier, max_over_ranks, shard_range, sampler, checks):
    q, n, n_lwe = BR_Q, 1 << BR_LOGN, BR_NLWE
    b0, b1 = shard_range(BR_TOTAL, world, rank)
    share = b1 - b0
    t10 = P.U32NttTable(BR_LOGN, q, device=local_rank)
    lv = P.ApproxSignedBasis(q, BR_LOGB, None, 32).decompose_length()
    rng = np.random.default_rng(0x5EED0005)                      # same key and test vector on every rank (replicated)
    bsk_h = rng.integers(0, q, n_lwe * 2 * lv * 2 * n, dtype=np.uint64).astype(np.uint32)
    tv_h = rng.integers(0, q, n, dtype=np.uint64).astype(np.uint32)
    lwe_all = rng.integers(0, 2 * n, (BR_TOTAL, n_lwe + 1), dtype=np.uint64).astype(np.uint32)
    lwe_h = np.ascontiguousarray(lwe_all[b0:b1])
    key = P.BootstrappingKey(t10, BR_LOGB, None, n_lwe, bsk_h)
    bsk = torch.from_numpy(bsk_h.view(np.int32)).cuda()
    lwe = torch.from_numpy(lwe_h.view(np.int32)).cuda()
    tv = torch.from_numpy(tv_h.view(np.int32)).cuda()
    acc = torch.empty((share, 2 * n), dtype=torch.int32, device="cuda")
    run = lambda: t10.blind_rotate_batch(BR_LOGB, None, bsk, n_lwe, lwe, tv, acc)
    run(); torch.cuda.synchronize()
    if rank == 0:  # parity at full depth, outside the timed region: first/last ciphertexts of this rank's share vs the oracle
        from oracle import oracle as O
        rows = [0, 1, share - 1]
        want = O.blind_rotate(O.U32NttTable(BR_LOGN, q), O.ApproxSignedBasis(q, BR_LOGB, None, 32), bsk_h, n_lwe,
                              np.ascontiguousarray(lwe_h[rows]), tv_h, batch=len(rows), threads=host_threads())
        ok = bool(np.array_equal(acc[rows].cpu().numpy().view(np.uint32), want))
        checks["blind_rotation_n512_full_depth"] = ok
        if not ok:
            raise SystemExit("bench.py: blind-rotation parity check FAILED - refusing to time a wrong kernel")
    reps = 5
    barrier()
    launches0 = P.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tb0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    barrier()
    tb1 = time.perf_counter()
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    launches = P.launch_count() - launches0
    value = BR_TOTAL / (ms * 1e-3)
    # end to end: host LWE samples in -> host LWE samples out (H2D, blind rotation, sample extraction, D2H inside the call)
    out_h = np.empty((share, n + 1), dtype=np.uint32)
    key.bootstrap_slices(lwe_h, tv_h, out_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        key.bootstrap_slices(lwe_h, tv_h, out_h)
    e2e = BR_TOTAL * 3 / max_over_ranks(time.perf_counter() - t0)
    if rank == 0:
        from oracle import oracle as O
        rows = [0, share - 1]
        accs = acc[rows].cpu().numpy().view(np.uint32)
        checks["bootstrap_slices_extract"] = bool(all(np.array_equal(out_h[r], O.extract_lwe(a, q, 32)) for r, a in zip(rows, accs)))
    res = {"metric": "bootstraps/s (LWE blind rotation, n=512, N=1024 RLWE, u32 q=132120577, base 2^7 l=3; BASELINE config 5)",
           "value": value, "unit": "bootstrap/s", "n_gpus": world, "scaling": "strong", "higher_is_better": True,
           "config": {"workload": f"{BR_TOTAL} ciphertexts in total, contiguous shards of {share} per GPU, one shared bootstrapping key "
                                  "(25.2 MB, replicated), accumulator resident in shared memory for all 512 CMux steps",
                      "batch_per_gpu": share},
           "ms_per_step": ms, "steps": reps, "gpu_launches": int(launches),
           "e2e": {"value": e2e, "unit": "bootstrap/s", "h2d_bytes_per_step": share * (n_lwe + 1) * 4, "d2h_bytes_per_step": share * (n + 1) * 4,
                   "api": "pfhe_bootstrap32_slices (host LWE mod 2N in -> blind rotation -> extract_lwe -> host LWE out; key handle resident)"}}
    if rank == 0:
        peak = modmul_peak(P, local_rank)
        per_gpu = (share / (ms * 1e-3))
        res["roofline"] = {"bound": "integer pipe (IMAD)", "achieved": per_gpu * BR_MODMULS, "peak": peak, "unit": "modmul/s",
                           "frac": per_gpu * BR_MODMULS / peak,
                           "algorithmic_modmuls_per_bootstrap": BR_MODMULS,
                           "peak_source": "measured live: pfhe_modmul_microbench kind 2 (bare u32 Shoup products = 1 IMAD.HI + 2 IMAD, 8 chains/thread)",
                           "kernel": "blind_rotate_n1024_kernel (lattice32.cu)",
                           "ncu": "profiles/r02_ncu_blind_rotate_u32_fast.txt: sm__pipe_fmaheavy_cycles_active 74.6 %, issue 61.9 %, "
                                  "2927 instructions per thread per CMux (r01: 4159)",
                           "clocks": sampler.window(tb0, tb1)}
    return res


def bench_extra(P, torch, np, table, data, g, local_rank, peak, extra, checks):
    from oracle import oracle as O
    batch = data.shape[0]

    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best * 1e-3

    def u64(x):
        return x.cpu().numpy().view(np.uint64)

    def roof(bytes_per_s):
        return {"GB/s": bytes_per_s / 1e9, "frac_of_hbm_peak": bytes_per_s / 1e9 / peak}

    ot = O.U64NttTable(LOG_N, Q)
    # inverse NTT (parity: round trip of the forward-transformed buffer rows)
    rows = [0, 1, batch - 1]
    before = u64(data[rows]).copy()
    dt_inv = timed(lambda: table.inverse_batch(data))   # 4 inverse transforms in total
    extra["intt_per_s_n4096_u64"] = batch / dt_inv
    want = before.copy()
    for _ in range(4):
        ot.inverse_batch(want, 1)
    checks["intt_n4096"] = bool(np.array_equal(u64(data[rows]), want))
    # fused polymul
    half = batch // 2
    a_, b_, c_ = data[:half], data[half:], torch.empty_like(data[:half])
    dt = timed(lambda: table.polymul_batch(a_, b_, c_))
    extra["polymul_per_s_n4096_u64"] = half / dt
    extra["polymul_n4096_u64_roofline"] = roof(half * 3 * N * 8 / dt)
    checks["polymul_n4096"] = bool(np.array_equal(u64(c_[:2]), ot.polymul_batch(u64(a_[:2]).copy(), u64(b_[:2]).copy(), 1)))
    # N = 8192 forward (second size of config C2)
    t13 = P.U64NttTable(13, Q, device=local_rank)
    d13 = data.view(batch // 2, 8192)
    r13 = u64(d13[:2]).copy()
    dt = timed(lambda: t13.forward_batch(d13))
    extra["ntt_fwd_per_s_n8192_u64"] = (batch // 2) / dt
    extra["ntt_fwd_n8192_u64_roofline"] = roof((batch // 2) * 2 * 8192 * 8 / dt)
    w13 = r13.copy()
    for _ in range(4):
        O.U64NttTable(13, Q).forward_batch(w13, 1)
    checks["ntt_fwd_n8192"] = bool(np.array_equal(u64(d13[:2]), w13))
    # N = 8192 fused product (C2 names the product at both sizes); inputs reduced mod q first (d13 holds transform outputs: already canonical)
    q13 = batch // 4
    a13, b13, c13 = d13[:q13], d13[q13:2 * q13], torch.empty_like(d13[:q13])
    dt = timed(lambda: t13.polymul_batch(a13, b13, c13))
    extra["polymul_per_s_n8192_u64"] = q13 / dt
    fp13 = 3 * (4096 * 13 * 8 + 2 * 8192 * 3 + 8192 * 4) + 8192 * 12
    extra["polymul_n8192_u64_roofline"] = dict(roof(q13 * 3 * 8192 * 8 / dt), fp64_instr_per_product=fp13, frac_of_fp64_pipe=q13 / dt * fp13 / FP64_PEAK)
    checks["polymul_n8192"] = bool(np.array_equal(u64(c13[:2]), O.U64NttTable(13, Q).polymul_batch(u64(a13[:2]).copy(), u64(b13[:2]).copy(), 1)))
    del c13
    # 60-bit prime (integer pipe)
    q60 = 1152921504606830593
    t60 = P.U64NttTable(LOG_N, q60, device=local_rank)
    d60 = torch.randint(0, q60, (batch // 4, N), dtype=torch.int64, device="cuda", generator=g)
    r60 = u64(d60[:2]).copy()
    dt = timed(lambda: t60.forward_batch(d60))
    extra["ntt_fwd_per_s_n4096_u64_q60"] = (batch // 4) / dt
    extra["ntt_fwd_n4096_u64_q60_roofline"] = roof((batch // 4) * BYTES_PER_NTT / dt)
    w60 = r60.copy()
    for _ in range(4):
        O.U64NttTable(LOG_N, q60).forward_batch(w60, 1)
    checks["ntt_fwd_n4096_q60"] = bool(np.array_equal(u64(d60[:2]), w60))
    del d60
    # C1: N=1024, one 32-bit prime (q = 132120577), round trip + fused product; batch 1024 is one launch of ~10 us, so the rate is
    # also taken at batch 2^20 (4 GiB of traffic) where the kernel, not the launch, is measured
    q32 = 132120577
    t10 = P.U32NttTable(10, q32, device=local_rank)
    o10 = O.U32NttTable(10, q32)
    for nb, tag in ((1024, "batch1024"), (1 << 20, "batch1M")):
        xa = torch.randint(0, q32, (nb, 1024), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
        xb = torch.randint(0, q32, (nb, 1024), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
        xc = torch.empty_like(xa)
        dt = timed(lambda: t10.polymul_batch(xa, xb, xc), reps=5)
        extra[f"c1_polymul_per_s_n1024_u32_{tag}"] = nb / dt
        if nb > 1024:
            pk32 = modmul_peak(P, local_rank)   # bare u32 Shoup products per second, measured live
            mm_poly = 3 * 512 * 10 + 1024       # three transforms of (N/2) log2 N butterflies + the pointwise products
            extra["c1_polymul_n1024_u32_roofline"] = dict(roof(nb * 3 * 1024 * 4 / dt), modmuls_per_product=mm_poly, modmul_peak=pk32,
                                                          frac_of_modmul_roof=nb / dt * mm_poly / pk32, bound="integer pipe (binding) / hbm")
            dt = timed(lambda: t10.forward_batch(xc), reps=5)
            extra["c1_ntt_fwd_n1024_u32_roofline"] = dict(roof(nb * 2 * 1024 * 4 / dt), ntt_per_s=nb / dt, frac_of_modmul_roof=nb / dt * 5120 / pk32)
        else:
            want = o10.polymul_batch(xa[:4].cpu().numpy().view(np.uint32).copy(), xb[:4].cpu().numpy().view(np.uint32).copy(), 1)
            rt = xa[:4].clone(); t10.forward_batch(rt); t10.inverse_batch(rt)
            checks["c1_polymul_and_round_trip_n1024_u32"] = bool(np.array_equal(xc[:4].cpu().numpy().view(np.uint32), want) and torch.equal(rt, xa[:4]))
        del xa, xb, xc
    # C4 external products: N=2048, k=1, base 2^7, batch 4096, shared key (A: u32 l=3; B: u64 l=7)
    for bits, q, tag in ((32, 132120577, "u32_l3"), (64, Q, "u64_l7")):
        tdt, ndt = (torch.int64, np.uint64) if bits == 64 else (torch.int32, np.uint32)
        t11 = (P.U64NttTable if bits == 64 else P.U32NttTable)(11, q, device=local_rank)
        o11 = (O.U64NttTable if bits == 64 else O.U32NttTable)(11, q)
        ob = O.ApproxSignedBasis(q, 7, None, bits); lv = ob.decompose_length()
        key = torch.randint(0, q, (2 * lv * 2 * 2048,), dtype=torch.int64, device="cuda", generator=g).to(tdt)
        cin = torch.randint(0, q, (4096, 2 * 2048), dtype=torch.int64, device="cuda", generator=g).to(tdt)
        cout = torch.empty_like(cin)
        dt = timed(lambda: t11.external_product_batch(1, 7, None, key, cin, cout, True))
        extra[f"external_products_per_s_n2048_{tag}"] = 4096 / dt
        if bits == 32:   # C4 instance A: 2l forward + 2 inverse transforms + 4lN multiply-accumulates = 116,736 modular multiplications
            mm = 2 * lv * 11264 + 4 * lv * 2048 + 2 * (11264 + 1024)
            pk = modmul_peak(P, local_rank)
            extra[f"external_product_n2048_{tag}_roofline"] = {"bound": "integer pipe (IMAD)", "modmuls_per_product": mm, "achieved": 4096 / dt * mm,
                                                               "peak": pk, "unit": "modmul/s", "frac": 4096 / dt * mm / pk,
                                                               "kernel": "external_product_u32_kernel<N=2048> (lattice32_ep.cu)"}
        else:            # C4 instance B: FP64-pipe bound (2l + 2 transforms of ~76 K FP64 instructions, 4lN products of ~7)
            fp = (2 * lv + 2) * (11264 * 8 + 2 * 2048 * 3) + 4 * lv * 2048 * 7
            extra[f"external_product_n2048_{tag}_roofline"] = {"bound": "fp64 pipe", "fp64_instr_per_product": fp, "frac": 4096 / dt * fp / FP64_PEAK,
                                                               "peak_instr_per_s": FP64_PEAK, "kernel": "external_product_kernel<F64LazyField> (lattice.cu)"}
        want = O.external_product_single(o11, ob, 1, key.cpu().numpy().view(ndt), cin[:2].cpu().numpy().view(ndt).copy(), to_coeff=True, batch=2, threads=1)
        checks[f"external_product_n2048_{tag}"] = bool(np.array_equal(cout[:2].cpu().numpy().view(ndt), want))
        # end to end through the host-slice shim (key uploaded per call, ciphertexts streamed)
        hin, hout = cin.cpu().numpy().view(ndt), np.empty((4096, 4096), dtype=ndt)
        kh = key.cpu().numpy().view(ndt)
        t11.external_product_slices(1, 7, None, kh, hin, hout, True)
        t0 = time.perf_counter(); t11.external_product_slices(1, 7, None, kh, hin, hout, True)
        extra[f"external_products_per_s_n2048_{tag}_e2e_pageable"] = 4096 / (time.perf_counter() - t0)
        del key, cin, cout
    # C3 RNS polynomial product: N=16384, 8 limbs of ~50-bit primes (q = 1 mod 2^15), fused per limb
    c3 = _c3_primes()
    dc = P.U64DcrtTable(14, c3, device=local_rank)
    nrns = 1024
    ra = torch.stack([torch.randint(0, m, (nrns, 16384), dtype=torch.int64, device="cuda", generator=g) for m in c3], dim=1).contiguous()
    rb, rc = ra.flip(0).contiguous(), torch.empty_like(ra)
    dt = timed(lambda: dc.polymul_batch(ra, rb, rc), reps=5)
    extra["rns_polymuls_per_s_n16384_l8_u64"] = nrns / dt
    fp_limb = 3 * (8192 * 14 * 8 + 2 * 16384 * 3 + 16384 * 4) + 16384 * 12   # FP64 instructions of one limb product (3 transforms + pointwise)
    extra["rns_polymul_n16384_l8_roofline"] = dict(roof(nrns * 3 * 8 * 16384 * 8 / dt), bound="fp64 pipe (binding) / hbm",
                                                   algorithmic_bytes_per_product=3 * 8 * 16384 * 8, fp64_instr_per_limb_product=fp_limb,
                                                   frac_of_fp64_pipe=nrns * 8 / dt * fp_limb / FP64_PEAK,
                                                   note="2-CTA thread-block cluster per polynomial, st.async exchange (ntt_cluster.cu); history in profiles/r02_large_n_experiments.md")
    want = np.stack([O.U64NttTable(14, m).polymul_batch(u64(ra[:1, i]).copy(), u64(rb[:1, i]).copy(), 1) for i, m in enumerate(c3)], axis=1)
    checks["rns_polymul_n16384_l8"] = bool(np.array_equal(u64(rc[:1]), want))
    fa = ra.clone()
    dt = timed(lambda: dc.forward_batch(fa))
    extra["dcrt_ntt_fwd_n16384_l8_roofline"] = roof(nrns * 8 * 2 * 16384 * 8 / dt)
    del ra, rb, rc, fa
    # multi-limb external product: N=2048, k=1, L=2 (100-bit Q), base 2^7 (l=14), batch 1024
    m2 = [Q, 1125899906629633]
    dc2 = P.U64DcrtTable(11, m2, device=local_rank)
    bb = P.BigUintApproxSignedBasis(P.RNSBase(m2, 64), 7, None)
    lv2 = bb.decompose_length()
    key2 = torch.stack([torch.randint(0, m, (2 * lv2 * 2, 2048), dtype=torch.int64, device="cuda", generator=g) for m in m2], dim=1).contiguous()
    cin2 = torch.stack([torch.randint(0, m, (1024 * 2, 2048), dtype=torch.int64, device="cuda", generator=g) for m in m2], dim=1).contiguous()
    cout2 = torch.empty_like(cin2)
    scratch = P.dcrt_external_product_batch(dc2, bb, 1, key2, cin2, cout2, True)
    extra["dcrt_external_products_per_s_n2048_L2_l14_u64"] = 1024 / timed(
        lambda: P.dcrt_external_product_batch(dc2, bb, 1, key2, cin2, cout2, True, scratch=scratch))
    orns = O.RNSBase(m2, 64)
    want = O.external_product(O.DcrtTable(11, m2, 64), orns, O.BigUintApproxSignedBasis(orns, 7, None), 1, u64(key2).reshape(-1),
                              u64(cin2[:2]).reshape(-1).copy(), to_coeff=True, batch=1, threads=1)
    checks["dcrt_external_product_n2048_L2"] = bool(np.array_equal(u64(cout2[:2]).reshape(-1), want.reshape(-1)))
    del key2, cin2, cout2, scratch
    # ---- streaming kernels (K4 / K6 / K9): algorithmic GB/s against the HBM peak, >= 1 GiB per operand --------------------------
    stream = {}
    nel = 1 << 27                                         # 128 Mi words = 1 GiB per u64 operand
    x = data.view(-1)[:nel]; y = data.view(-1)[nel:2 * nel]
    o = torch.empty_like(x)
    bm = P.BarrettModulus(Q, 64)
    for name, fn, words in (("reduce_mul_slice_to", lambda: bm.reduce_mul_slice_to(x, y, o), 3),
                            ("reduce_add_mul_slice_assign", lambda: bm.reduce_add_mul_slice_assign(o, x, y), 4),
                            ("reduce_add_slice_to", lambda: bm.reduce_add_slice_to(x, y, o), 3),
                            ("factor_mul_slice_to", lambda: bm.factor_mul_slice_to(12345678901, x, o), 2)):
        dt = timed(fn)
        stream[name] = roof(nel * 8 * words / dt)
    checks["slice_mul"] = True
    bm.reduce_mul_slice_to(x, y, o)
    xs, ys = u64(x[:4096]), u64(y[:4096])
    checks["slice_mul"] = bool(np.array_equal(u64(o[:4096]), O.BarrettModulus(Q, 64).reduce_mul_slice_to(xs.copy(), ys.copy())))
    # gadget decomposition (1 + l words per coefficient) and the fused multi-word gadget (L in, l*L out)
    lvq = P.ApproxSignedBasis(Q, 7, None, 64)
    nd = 1 << 24
    dig = torch.empty((lvq.decompose_length(), nd), dtype=torch.int64, device="cuda")
    dt = timed(lambda: lvq.decompose_batch(x[:nd], dig))
    stream["decompose_l7"] = roof(nd * 8 * (1 + lvq.decompose_length()) / dt)
    rns = P.RNSBase(m2, 64)
    bb2 = P.BigUintApproxSignedBasis(rns, 7, None)
    polys = 2048
    res_in = torch.stack([torch.randint(0, m, (polys, 2048), dtype=torch.int64, device="cuda", generator=g) for m in m2], dim=1).contiguous()
    digs = torch.empty((polys, bb2.decompose_length(), 2, 2048), dtype=torch.int64, device="cuda")
    dt = timed(lambda: bb2.gadget_decompose_batch(res_in, digs, 2048))
    stream["rns_gadget_L2_l14"] = roof(polys * 2048 * 8 * (2 + 2 * bb2.decompose_length()) / dt)
    del res_in, digs
    polys = 16384                                          # 32 Mi coefficients: 256 MiB per limb, beyond L2
    cnt = polys * 2048
    flat = torch.stack([torch.randint(0, m, (cnt,), dtype=torch.int64, device="cuda", generator=g) for m in m2], dim=0).contiguous()
    big = torch.empty((cnt, rns.big_uint_value_len()), dtype=torch.int64, device="cuda")
    dt = timed(lambda: rns.compose_multiple_values_to(flat, big))
    stream["rns_compose_L2"] = roof(cnt * 8 * (2 + rns.big_uint_value_len()) / dt)
    back = torch.empty_like(flat)
    dt = timed(lambda: rns.decompose_big_uint_values_to(big, back))
    stream["rns_decompose_L2"] = roof(cnt * 8 * (2 + rns.big_uint_value_len()) / dt)
    checks["rns_compose_decompose_round_trip"] = bool(torch.equal(back, flat)) and bool(np.array_equal(
        u64(big[:4096]).reshape(-1), O.RNSBase(m2, 64).compose_multiple_values_to(u64(flat[:, :4096]).copy().reshape(-1), 4096)))
    del flat, big, back
    bc_in = [137438822401, 137438814209, 137438773249]
    bc = P.BaseConverter(bc_in, [Q, 1125899906629633], 64)
    cin3 = torch.stack([torch.randint(0, m, (polys, 2048), dtype=torch.int64, device="cuda", generator=g) for m in bc.in_moduli], dim=1).contiguous()
    cout3 = torch.empty((polys, 2, 2048), dtype=torch.int64, device="cuda")
    dt = timed(lambda: bc.fast_convert_array(cin3, cout3, 2048))
    stream["baseconv_fast_3_to_2"] = roof(cnt * 8 * 5 / dt)
    obc = O.BaseConverter(bc_in, [Q, 1125899906629633], 64)
    checks["baseconv_fast_3_to_2"] = bool(np.array_equal(u64(cout3[polys - 1]).reshape(-1), obc.fast_convert_array(u64(cin3[polys - 1]).copy().reshape(-1), 2048)))
    bc1 = P.BaseConverter(bc_in, [Q], 64)
    cex = cout3.view(-1)[:cnt]
    dt = timed(lambda: bc1.exact_convert_array(cin3, cex, 2048))
    stream["baseconv_exact_3_to_1"] = roof(cnt * 8 * 4 / dt)
    del cin3, cout3
    extra["streaming_kernels"] = stream


def run_capi_driver(args):
    """ONE process, --gpus devices, no torch.distributed: the headline host-slice workload and the bootstrap workload through
    pfhe_multi_* (one host thread + stream set per device inside the library)."""
    import numpy as np
    import primus_fhe_b200 as P
    ndev = args.gpus
    devs = list(range(ndev))
    if P.device_count() < ndev:
        raise SystemExit(f"--driver capi --gpus {ndev}: only {P.device_count()} devices visible")
    import torch
    mt = P.MultiNttTable(LOG_N, Q, devs, 64)
    per = args.batch // 4
    host = torch.empty((per * ndev, N), dtype=torch.int64).pin_memory()
    rng = np.random.default_rng(0x5EED0002)
    host.numpy().view(np.uint64)[:] = rng.integers(0, Q, (per * ndev, N), dtype=np.uint64)
    mt.transform_slices(host)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        mt.transform_slices(host)
    ntt_e2e = per * ndev * reps / (time.perf_counter() - t0)
    q, n, n_lwe = BR_Q, 1 << BR_LOGN, BR_NLWE
    m32 = P.MultiNttTable(BR_LOGN, q, devs, 32)
    lv = P.ApproxSignedBasis(q, BR_LOGB, None, 32).decompose_length()
    rng = np.random.default_rng(0x5EED0005)
    bsk = rng.integers(0, q, n_lwe * 2 * lv * 2 * n, dtype=np.uint64).astype(np.uint32)
    tv = rng.integers(0, q, n, dtype=np.uint64).astype(np.uint32)
    lwe = rng.integers(0, 2 * n, (BR_TOTAL, n_lwe + 1), dtype=np.uint64).astype(np.uint32)
    keys = m32.bootstrapping_keys(BR_LOGB, None, n_lwe, bsk)
    out = np.empty((BR_TOTAL, n + 1), dtype=np.uint32)
    m32.bootstrap_slices(keys, lwe, tv, out)
    t0 = time.perf_counter()
    for _ in range(reps):
        m32.bootstrap_slices(keys, lwe, tv, out)
    bs = BR_TOTAL * reps / (time.perf_counter() - t0)
    print(json.dumps({"driver": "capi (pfhe_multi_*: one process, one host thread per device, no torch.distributed)", "n_gpus": ndev,
                      "e2e_ntt_per_s_n4096": ntt_e2e, "ntt_batch_per_gpu": per, "e2e_bootstraps_per_s": bs, "bootstrap_total": BR_TOTAL,
                      "gpu_launches": P.launch_count()}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
