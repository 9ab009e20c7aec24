#!/usr/bin/env python
"""bench.py -- headline benchmark of the polynomial-ring hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], the configuration the metric "NTTs/s at N=4096" is quoted on):
batched forward negacyclic NTT, N = 4096, 64-bit NTT-friendly prime q = 1125899906826241 (the reference's
own bench prime, primus_ntt/benches/bench_u64.rs:8), batch 65536 polynomials (2 GiB) per GPU.
A "step" is one pass of the hot path over the whole batch (one kernel launch, in place).

  value  = whole-job NTTs/s with the batch resident in HBM (CUDA events, max over ranks)
  e2e    = the same metric through the reference-facing C-ABI host-slice call
           (pfhe_ntt64_transform_slices: pinned HOST buffer -> H2D -> kernel -> D2H, all inside the timed region)
  roofline = algorithmic bytes (2*N*8 per NTT) / kernel time against the measured HBM copy peak
  cpu_baseline = the CPU oracle (restated reference scalar path, OpenMP over the batch) on this box's cores
  extra  = secondary numbers of the same path (INTT, fused polymul, external product, blind rotation = bootstraps/s)

Multi-GPU: one process per GPU (torchrun), the batch is sharded (independent polynomials, no collective on
the data path; NCCL only for the barrier / max-over-ranks of the timing) -> "scaling": "weak".

`--impl reference` times the reference arm: the reference's CPU algorithm (oracle port, all host threads) on a
bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
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


def _bind_to_gpu_numa_node(gpu_index: int):
    """Pin this process (and therefore the first-touch placement of its pinned host buffers) to the NUMA node the GPU
    hangs off: the e2e path moves 4 GiB per step over PCIe, cross-socket traffic halves it at 8 ranks."""
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
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        return None
    return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons through NVML (in-process thread, ~2 ms period) while the timed
    region runs; falls back to `nvidia-smi -lms` when NVML is unavailable."""
    REASONS = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.stop_flag, self.thread, self.proc, self.mx = gpu_index, [], False, None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical GPUs; honour CUDA_VISIBLE_DEVICES when it lists indices
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
                        self.rows.append((time.perf_counter(), float(clk), int(rs)))
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self, t_begin=None, t_end=None):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        rows = self.rows
        inside = [r for r in rows if t_begin is not None and t_begin <= r[0] <= t_end]
        note = "sampled inside the timed region"
        if len(inside) < 3:
            inside, note = rows, "timed region shorter than 3 samples: all samples of this run (warm-up + timed + e2e) used"
        sm = [r[1] for r in inside]
        reasons = set()
        for r in inside:
            for name, bit in self.REASONS.items():
                if r[2] & bit:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(reasons),
                "samples": len(sm), "note": note}


def cpu_reference(sample_batch: int, reps: int):
    """The reference's CPU path (oracle port): forward NTT over `sample_batch` polys, all host threads."""
    import numpy as np
    from oracle import oracle as O
    t = O.U64NttTable(LOG_N, Q)
    threads = O.max_threads()
    rng = np.random.default_rng(0x5EED0002)
    x = rng.integers(0, Q, (sample_batch, N), dtype=np.uint64)
    t.forward_batch(x[:min(256, sample_batch)].copy(), threads)  # warm
    best = 1e30
    for _ in range(reps):
        y = x.copy()
        t0 = time.perf_counter()
        t.forward_batch(y, threads)
        best = min(best, time.perf_counter() - t0)
    return sample_batch / best, threads, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = 16384
    times = []
    import numpy as np
    from oracle import oracle as O
    t = O.U64NttTable(LOG_N, Q)
    threads = O.max_threads()
    rng = np.random.default_rng(0x5EED0002)
    x = rng.integers(0, Q, (sample, N), dtype=np.uint64)
    for i in range(args.warmup + args.steps):
        y = x.copy()
        t0 = time.perf_counter()
        t.forward_batch(y, threads)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": dict(CONFIG, reference_sample=f"{sample} polynomials per step (bounded sample of the 65536-poly batch)"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} NTTs/step x {args.steps} steps, OpenMP over the batch; the Rust reference cannot be "
                                       "built here (no cargo) so this is the C restatement of its scalar Harvey path"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary (extra) measurements")
    ap.add_argument("--batch", type=int, default=BATCH)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import primus_fhe_b200 as P

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_node = _bind_to_gpu_numa_node(local_rank) if world > 1 else None
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

    batch = args.batch
    table = P.U64NttTable(LOG_N, Q, device=local_rank)
    g = torch.Generator(device="cuda"); g.manual_seed(0x5EED0002 + rank)
    data = torch.randint(0, Q, (batch, N), dtype=torch.int64, device="cuda", generator=g)

    # ---- device-resident throughput (value) --------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
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
    t_ms = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms_max = float(t_ms.item())
    value = world * batch * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the C-ABI host-slice call (pinned host buffers) ----------------------------
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
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_batch * e2e_steps / float(t_e.item())
    del host
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel ----------------------------------------------------------
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
                "peak_source": peak_src, "kernel": "ntt_tma_kernel<F64LazyField,N=4096> forward (lazy-fold FP64 butterflies, TMA tensor store)", "algorithmic_bytes_per_launch": batch * BYTES_PER_NTT,
                "avg_launch_ms": avg_kernel_ms,
                "best_launch_ms": min(kernel_ms), "frac_best_launch": batch * BYTES_PER_NTT / (min(kernel_ms) * 1e-3) / 1e9 / peak,
                "note": "achieved/frac use the AVERAGE launch over the timed region (the board reaches its power cap after ~0.1 s of this "
                        "FP64-heavy kernel: see clocks.sm_mhz / reasons); peak is the burst copy bandwidth; best_launch is the fastest "
                        "single launch of the same region",
                "modmul": {"butterflies_per_s": batch * MODMULS_PER_NTT / (avg_kernel_ms * 1e-3),
                           "fp64_instr_per_ntt": FP64_PER_NTT, "fp64_pipe_peak_instr_per_s": FP64_PEAK,
                           "frac_of_fp64_pipe": batch * FP64_PER_NTT / (avg_kernel_ms * 1e-3) / FP64_PEAK,
                           "note": "secondary (binding) bound: q < 2^50 runs on the FP64 pipe, 944 FP64 instructions per thread x 256 "
                                   "threads per NTT against the measured DFMA rate (profiles/r01_ubench_pipes.log); see DESIGN.md section 4"}}

    # ---- CPU baseline (bounded sample) ---------------------------------------------------------------------
    cpu_value, cores, cpu_s = cpu_reference(8192, 3)
    cpu_baseline = {"value": cpu_value, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"8192 NTTs (1/8 of the batch), best of 3, {cpu_s:.3f} s; C restatement of the reference's scalar "
                              "Harvey NTT (the Rust reference cannot be built in this image)"}

    # ---- secondary measurements (same hot path; a few launches each) ------------------------------------------
    extra = {}
    if not args.no_extra:
        def timed(fn, reps=3):
            fn(); torch.cuda.synchronize()
            best = 1e30
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            return best * 1e-3
        try:
            extra["intt_per_s_n4096_u64"] = batch / timed(lambda: table.inverse_batch(data))
            half = batch // 2
            a_, b_, c_ = data[:half], data[half:], torch.empty_like(data[:half])
            extra["polymul_per_s_n4096_u64"] = half / timed(lambda: table.polymul_batch(a_, b_, c_))
            # C4-B external product: N=2048, k=1, base 2^7 (l=7), batch 4096, shared key
            t11 = P.U64NttTable(11, Q, device=local_rank)
            lv = P.ApproxSignedBasis(Q, 7, None, 64).decompose_length()
            key = torch.randint(0, Q, (2 * lv * 2 * 2048,), dtype=torch.int64, device="cuda", generator=g)
            cin = torch.randint(0, Q, (4096, 2 * 2048), dtype=torch.int64, device="cuda", generator=g)
            cout = torch.empty_like(cin)
            extra["external_products_per_s_n2048_u64_l7"] = 4096 / timed(lambda: t11.external_product_batch(1, 7, None, key, cin, cout, True))
            # C5 blind rotation: n=512, N=1024, u32 q=132120577, base 2^7 (l=3); per-GPU share of the 10k batch
            q32, nl = 132120577, 512
            t10 = P.U32NttTable(10, q32, device=local_rank)
            lv3 = P.ApproxSignedBasis(q32, 7, None, 32).decompose_length()
            bsk = torch.randint(0, q32, (nl * 2 * lv3 * 2 * 1024,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
            nb = 1250
            lwe = torch.randint(0, 2048, (nb, nl + 1), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
            tv = torch.randint(0, q32, (1024,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
            acc = torch.empty((nb, 2048), dtype=torch.int32, device="cuda")
            extra["bootstraps_per_s_blind_rotation_n512_N1024_u32"] = nb / timed(lambda: t10.blind_rotate_batch(7, None, bsk, nl, lwe, tv, acc), reps=2)
            del bsk, lwe, acc, key, cin, cout
            # C3 RNS polynomial product: N=16384, 8 limbs of ~50-bit primes (q = 1 mod 2^15), fused per limb
            c3 = _c3_primes()
            dc = P.U64DcrtTable(14, c3, device=local_rank)
            nrns = 256
            ra = torch.stack([torch.randint(0, m, (nrns, 16384), dtype=torch.int64, device="cuda", generator=g) for m in c3], dim=1).contiguous()
            rb, rc = ra.flip(0).contiguous(), torch.empty_like(ra)
            extra["rns_polymuls_per_s_n16384_l8_u64"] = nrns / timed(lambda: dc.polymul_batch(ra, rb, rc))
            del ra, rb, rc
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
        except Exception as ex:  # secondary numbers must never hide the headline
            extra["error"] = repr(ex)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": dict(CONFIG, batch_per_gpu=batch, n_gpus=world),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_batch * N * 8, "d2h_bytes_per_step": e2e_batch * N * 8,
                    "api": "pfhe_ntt64_transform_slices (host-slice shim of NttTable::transform_slice, pinned host memory)",
                    "numa_node_rank0": numa_node,
                    "steps": e2e_steps},
            "gpu_launches": int(launches), "clocks": clocks, "extra": extra}
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
