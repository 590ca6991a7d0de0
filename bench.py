#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 ring-arithmetic hot path (BASELINE.json configs[1]).

A "step" is one pass of the hot path over one batch of synthetic input: `batch` independent negacyclic
R_q polymuls c = intt(ntt(a) . ntt(b)) at N=1024, q=65537 (the reference's NTT prime; ring_nq::mul,
arith/src/ring_nq.rs:586-607).

  value        whole-job polymul/s, operands resident in HBM as u64 words (the layout of SURVEY 8b)
  roofline     the polymul kernel against the measured HBM peak (live CUDA events), plus -- inside the same key,
               because the driver keeps it -- `int` (integer-pipe roofline), `sustained` (>= 2 s loop), `sweep`
               (polymul at N=2^10..2^14, q=65537 and a 62-bit prime, u64 and u32 device formats, each against the
               slower of its HBM and modmul rooflines), `bootstrap` (TFHE bootstraps/s against the MEASURED int8
               tensor peak, burst and sustained), `bfv`, `extprod`
  e2e          the same metric through the C ABI with pinned HOST buffers, H2D + D2H inside the timed region, on the
               bit-packed wire (17 bits per coefficient); the u32 and u64 wires and a PCIe copy microbenchmark beside it
  cpu_baseline the oracle port (the reference's CPU algorithm) timed on this box's host cores

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
Under torchrun (N>1) every rank runs the same per-GPU batch (weak scaling, no data-path collective).
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

Q = 65537
N = 1024
BATCH = 65536  # 3 * 65536 * 8 KiB = 1.5 GiB per step: far larger than the 126 MB L2
WIRE_BITS = 17  # ceil(log2 q): the bit-packed wire of the end-to-end leg
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of this exact workload (ncu --set full capture)
# Twiddle products the Fermat32 kernel executes per N=1024 polymul (radix-4: three products and a shift per four
# butterflies, and the first layer of a forward transform is shifts only): forward 3 radix-4 layers x 256 blocks x 3 +
# 2 radix-2 stages x 512 = 3328, twice; inverse 4 x 256 x 3 + 512 + the n^-1 stage's 1024 = 4608; pointwise 1024.
# SURVEY 8d's algorithmic count for the same polymul is 17408.
MODMUL_EXECUTED = 2 * 3328 + 4608 + 1024
NCU_TRAFFIC_BYTES = 1073800000 + 498943744
NCU_TRAFFIC_SOURCE = "profiles/r2_polymul_n1024_q65537_u64_fermat32_ncu_full_d.csv (ncu --set full of tools/prof.py polymul 10 65537 65536)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def workload_config(batch, world=1, extra=None):
    """The `config` object, identical in both arms."""
    cfg = {
        "workload": f"BASELINE configs[1]: batched Rq negacyclic NTT polymul, N={N}, q={Q}, batch {batch} per GPU",
        "batch_per_gpu": batch, "n": N, "q": Q,
        "l2_policy": f"inputs+outputs {3 * N * 8 * batch / 2**20:.0f} MiB per step, larger than the 126 MB L2",
        "parallelism": f"independent polynomials sharded over {world} GPU(s), no collective",
    }
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    NAMES = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
    }

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {
            "sm_mhz": statistics.median(self.samples),
            "sm_min_mhz": min(self.samples),
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def cpu_polymul_baseline(target_s: float = 12.0):
    """Oracle port (oracle/fhe_oracle.c: u128 % q butterflies, arith/src/ntt.rs + ring_nq.rs) on all host cores."""
    import oracle

    cores = os.cpu_count() or 1
    a = oracle.uniform(1, (cores * 64, N), Q)
    b = oracle.uniform(2, (cores * 64, N), Q)
    oracle.rq_mul_batch(Q, N, a, b, threads=cores)  # thread pool and pages warm
    t0 = time.perf_counter()
    oracle.rq_mul_batch(Q, N, a, b, threads=cores)
    dt = time.perf_counter() - t0
    rate = a.shape[0] / dt
    sample = int(max(cores * 64, min(rate * target_s, 4_000_000)))
    a = oracle.uniform(3, (sample, N), Q)
    b = oracle.uniform(4, (sample, N), Q)
    t0 = time.perf_counter()
    c = oracle.rq_mul_batch(Q, N, a, b, threads=cores)
    dt = time.perf_counter() - t0
    return {
        "value": sample / dt, "unit": "polymul/s", "cores": cores, "kind": "port",
        "sample": f"{sample} polymuls N={N} q={Q} (oracle C port built -O3 -march=native, OpenMP over the batch), {dt:.2f} s",
    }, (a, b, c)


def run_reference(args, emit):
    """--impl reference: the reference's CPU algorithm (oracle port; no Rust toolchain exists here) on all host
    cores, same metric/config.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    cores = os.cpu_count() or 1
    # a step = a bounded sample of the workload: ~0.35 s of CPU work, so warm-up + timed steps last seconds, not milliseconds
    oracle.rq_mul_batch(Q, N, oracle.uniform(1, (cores * 64, N), Q), oracle.uniform(2, (cores * 64, N), Q), threads=cores)
    a = oracle.uniform(3, (cores * 256, N), Q)
    b = oracle.uniform(4, (cores * 256, N), Q)
    t0 = time.perf_counter()
    oracle.rq_mul_batch(Q, N, a, b, threads=cores)
    rate = a.shape[0] / (time.perf_counter() - t0)
    per_step = int(min(BATCH, max(cores * 256, rate * 0.35)))
    a = oracle.uniform(5, (per_step, N), Q)
    b = oracle.uniform(6, (per_step, N), Q)
    for _ in range(args.warmup):
        oracle.rq_mul_batch(Q, N, a, b, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.rq_mul_batch(Q, N, a, b, threads=cores)
    dt = time.perf_counter() - t0
    v = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "NTT polymul/s", "value": v, "unit": "polymul/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(BATCH, args.gpus),
        "cpu_baseline": {"value": v, "unit": "polymul/s", "cores": cores, "kind": "port",
                         "sample": f"each step = {per_step} polymuls of the {BATCH}-polymul batch (throughput is per polymul), "
                                   f"oracle C port of arith/src/ntt.rs + ring_nq.rs built -O3 -march=native, {dt:.2f} s timed"},
        "e2e": {"value": v, "unit": "polymul/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def measure_pcie(torch, dev, mib=256, reps=8, sync=None):
    """Pinned-memory copy rates of this rank: H2D alone, D2H alone, both at once (GB/s).  `sync` (a barrier) is called in
    front of every phase so that all ranks copy at the same time: the per-rank numbers then show what the host gives
    each GPU when N of them transfer concurrently."""
    n = mib << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        if sync is not None:
            sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        s1.synchronize()
        s2.synchronize()
        e1.record()
        torch.cuda.synchronize()
        return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d()
        d2h()

    # events on the default stream bracket work of the side streams: order them explicitly
    def wrap(fn):
        def g():
            s1.wait_stream(torch.cuda.current_stream())
            s2.wait_stream(torch.cuda.current_stream())
            fn()
            torch.cuda.current_stream().wait_stream(s1)
            torch.cuda.current_stream().wait_stream(s2)
        return g

    return {"h2d_gbs": timed(wrap(h2d)), "d2h_gbs": timed(wrap(d2h)), "bidir_each_gbs": timed(wrap(both)), "mib": mib}


def run_for(fn, seconds, torch, chunk=32):
    """Call fn back to back for at least `seconds` of device time; returns (calls, elapsed_s)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    calls, elapsed = 0, 0.0
    while elapsed < seconds:
        e0.record()
        for _ in range(chunk):
            fn()
        e1.record()
        torch.cuda.synchronize()
        elapsed += e0.elapsed_time(e1) * 1e-3
        calls += chunk
    return calls, elapsed


def measure_bootstrap(fhe, torch, dist, dev, rank, world, local, quick):
    """Second headline metric (BASELINE metric: "TFHE bootstraps/s"): bootstrapping as the reference executes
    it (tfhe/src/tlwe.rs:150-161; n=1024, k=1, l=64), TLWE inputs sharded over the ranks, the 537 MB
    key-switching key and the table broadcast ONCE from rank 0 over NCCL, no per-op collective."""
    from fhe_study_b200.dist import broadcast_key

    n, k, kn, l = 1024, 1, 1024, 64
    batch = 2048 if quick else 8192
    steps = 2 if quick else 5
    g = torch.Generator(device=dev).manual_seed(99)
    if rank == 0:
        ksk = torch.randint(-(2**63), 2**63 - 1, (kn * l * (kn + 1),), dtype=torch.int64, device=dev, generator=g)
        table = torch.randint(-(2**63), 2**63 - 1, ((k + 1) * n,), dtype=torch.int64, device=dev, generator=g)
    else:
        ksk = torch.empty((kn * l * (kn + 1),), dtype=torch.int64, device=dev)
        table = torch.empty(((k + 1) * n,), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    broadcast_key(ksk)
    broadcast_key(table)
    b1.record()
    torch.cuda.synchronize()
    bcast_ms = b0.elapsed_time(b1)
    K = fhe.Ksk(kn, kn, l, ksk)
    del ksk
    cts = torch.randint(-(2**63), 2**63 - 1, (batch, kn + 1), dtype=torch.int64, device=dev,
                        generator=torch.Generator(device=dev).manual_seed(7 + rank))
    out = torch.empty_like(cts)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ms(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(3):
        fhe.bootstrap(n, k, K, table, cts, kn, out=out)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fhe.bootstrap(n, k, K, table, cts, kn, out=out)
    e1.record()
    barrier()
    ms = max_ms(e0.elapsed_time(e1))
    clocks_burst = sampler.stop()
    # sustained: the same call back to back for >= 2 s (the SM clock settles under the 1 kW cap)
    sus = {}
    if not quick:
        sampler = ClockSampler(local, period_s=0.02)
        sampler.start()
        calls, el = run_for(lambda: fhe.bootstrap(n, k, K, table, cts, kn, out=out), 2.0, torch, chunk=16)
        barrier()
        el = max_ms(el * 1e3) * 1e-3
        sus = {"value": world * batch * calls / el, "unit": "bootstraps/s", "seconds": el, "calls": calls, "batch_per_gpu": batch,
               "clocks": sampler.stop()}
        # upper end of BASELINE configs[4] (8k-64k ciphertexts): one GPU takes the whole 65536
        big = 65536
        cb = torch.randint(-(2**63), 2**63 - 1, (big, kn + 1), dtype=torch.int64, device=dev,
                           generator=torch.Generator(device=dev).manual_seed(70 + rank))
        ob = torch.empty_like(cb)
        fhe.bootstrap(n, k, K, table, cb, kn, out=ob)
        barrier()
        e0.record()
        for _ in range(3):
            fhe.bootstrap(n, k, K, table, cb, kn, out=ob)
        e1.record()
        barrier()
        sus["batch_65536_per_gpu"] = {"value": world * big * 3 / (max_ms(e0.elapsed_time(e1)) * 1e-3), "unit": "bootstraps/s"}
        del cb, ob
    # end to end through fhe_bootstrap with pinned HOST buffers (chunked H2D / compute / D2H overlap inside the call)
    hct = torch.empty(cts.shape, dtype=torch.int64).pin_memory()
    hout = torch.empty(cts.shape, dtype=torch.int64).pin_memory()
    hct.copy_(cts)
    fhe.bootstrap(n, k, K, table, hct, kn, out=hout)
    barrier()
    e0.record()
    for _ in range(steps):
        fhe.bootstrap(n, k, K, table, hct, kn, out=hout)
    e1.record()
    barrier()
    ms_e2e = max_ms(e0.elapsed_time(e1))
    e2e_ok = bool(torch.equal(hout.to(dev), out))
    value = world * batch * steps / (ms * 1e-3)
    mac = kn * l * (kn + 1)
    res = {
        "metric": "TFHE bootstraps/s (as executed: mod_switch + rotate + sample_extract + key_switch)",
        "value": value, "unit": "bootstraps/s", "n_gpus": world,
        "batch_per_gpu": batch, "steps": steps, "ms_per_step": ms / steps, "scaling": "weak",
        "params": {"n": n, "k": k, "l": l, "ksk_bytes": mac * 8},
        "key_broadcast_ms": bcast_ms,
        "u64_mac_per_s": value * mac,
        "int8_pops": value * mac * 16 / 1e15,  # every u64 MAC = 8 byte-plane int8 MACs = 16 int8 ops
        "clocks": clocks_burst,
        "sustained": sus,
        "e2e": {"value": world * batch * steps / (ms_e2e * 1e-3), "unit": "bootstraps/s",
                "h2d_bytes_per_step": batch * (kn + 1) * 8, "d2h_bytes_per_step": batch * (kn + 1) * 8,
                "matches_device_result": e2e_ok},
    }
    if rank == 0:  # measured tensor peaks of this GPU (no nominal figure involved)
        try:
            pk_burst, pk_sus = fhe.int_peak(3), (fhe.int_peak(4) if not quick else None)
            per_gpu = res["int8_pops"] / world
            res["roofline"] = {
                "bound": "tensor (int8)", "achieved": per_gpu, "unit": "POP/s per GPU", "peak": pk_burst / 1e15,
                "frac": per_gpu / (pk_burst / 1e15), "peak_source": "measured: fhe_int_peak(3), tcgen05.mma.kind::i8 loop, burst, this run",
                "peak_sustained": (pk_sus / 1e15) if pk_sus else None,
                "sustained_frac": (sus["value"] / world * mac * 16 / pk_sus) if (pk_sus and sus) else None,
                "nominal_dense_int8_pops": 4.5,
            }
        except Exception as ex:  # pragma: no cover
            res["roofline"] = {"error": repr(ex)}
    return res


def bind_to_gpu_numa(index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the end-to-end
    legs are allocated on (and copied from) the NUMA node the GPU hangs off.  Best effort; returns a note."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to gpu {index}"
    except Exception as ex:  # no NVML, no permission: keep the inherited affinity
        return f"unbound ({type(ex).__name__})"
    return "unbound"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # the contract is ONE JSON line on stdout: anything libraries print (e.g. NCCL's version banner) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        return run_reference(args, emit)

    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    numa_note = bind_to_gpu_numa(local) if world > 1 else "single rank: inherited affinity"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import fhe_study_b200 as fhe

    fhe.use_torch_stream()
    dev = torch.device("cuda", local)
    batch = args.batch
    quick = args.steps < 20
    plan = fhe.NttPlan(Q, N)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    a = torch.randint(0, Q, (batch, N), dtype=torch.int64, device=dev, generator=g)
    b = torch.randint(0, Q, (batch, N), dtype=torch.int64, device=dev, generator=g)
    c = torch.empty_like(a)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_list(x):
        if world == 1:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return [float(o.item()) for o in outs]

    # ---- device-resident throughput ("value") + per-launch kernel time ("roofline") ------------------
    for _ in range(args.warmup):
        plan.mul(a, b, out=c)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = fhe.launch_count()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    # The step is ONE launch of the polymul kernel, so the timed region holds nothing but the K launches back to back
    # on the stream: an event pair around every launch (as this region had before) costs 2-3 us of stream bubbles per
    # step, 1 % of a 0.26 ms kernel.  kernel_ms = region / launches (launch gaps included: conservative);
    # the per-launch event pairs are a second pass, reported beside it as kernel_ms_event_pairs.
    t_start.record()
    for _ in range(args.steps):
        plan.mul(a, b, out=c)
    t_end.record()
    barrier()
    launches = fhe.launch_count() - l0
    total_ms = t_start.elapsed_time(t_end)
    kern_ms = total_ms / max(1, launches)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 50))]
    for s, e in ev:
        s.record()
        plan.mul(a, b, out=c)
        e.record()
    barrier()
    kern_ms_pairs = statistics.mean(s.elapsed_time(e) for s, e in ev)
    total_ms_max = max_over_ranks(total_ms)
    value = world * batch * args.steps / (total_ms_max * 1e-3)
    clocks = sampler.stop()
    # integer-pipe peak measured now, in the same clock state as the timed region (register-only microbenchmark)
    modmul_peak = fhe.int_peak(1) if rank == 0 else None

    # ---- the same step sustained for >= 2 s (the timed region above is a burst of steps x 0.3 ms) ------------
    sustained = None
    if not quick:
        sampler = ClockSampler(local, period_s=0.02)
        sampler.start()
        calls, el = run_for(lambda: plan.mul(a, b, out=c), 2.0, torch, chunk=256)
        barrier()
        el = max_over_ranks(el)
        sustained = {"value": world * batch * calls / el, "unit": "polymul/s", "seconds": el, "calls": calls,
                     "hbm_frac": (3 * N * 8 * batch * calls / el / 1e9) / peaks()[0], "clocks": sampler.stop()}

    # ---- the packed 32-bit DEVICE format (q <= 2^32): same kernel instantiated with u32 loads and stores ---------
    a32, b32 = a.to(torch.int32), b.to(torch.int32)
    c32 = torch.empty_like(a32)
    for _ in range(3):
        plan.mul_u32(a32, b32, out=c32)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        plan.mul_u32(a32, b32, out=c32)
    e1.record()
    barrier()
    ms_u32 = e0.elapsed_time(e1) / args.steps
    u32_same = bool(torch.equal(c32.to(torch.int64), c))
    value_u32 = world * batch / (max_over_ranks(ms_u32) * 1e-3)
    del a32, b32, c32

    # ---- end to end through the C ABI with pinned host buffers ---------------------------------------
    # Three wire formats of the same call: bit-packed (fhe_rq_mul_packed, 17 bits per coefficient: the headline -- the step
    # is PCIe-bound, so bytes on the wire are what counts), u32 words (fhe_rq_mul_u32) and u64 words (fhe_rq_mul).  The Rust
    # shim gathers Vec<Zq>.v (16-byte AoS) into any of them in one pass over the operands.
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_run(ha, hb, hc, call, check):
        call(ha, hb, hc)  # warm the staging pool
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e2e_steps):
            call(ha, hb, hc)  # H2D(a,b) -> kernel -> D2H(c); returns when c is on the host
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        return world * batch * e2e_steps / (ms * 1e-3), check(hc)

    def pinned(dtype, cols, src=None):
        t = torch.empty((batch, cols), dtype=dtype).pin_memory()
        if src is not None:
            t.copy_(src)
        return t

    pcie = measure_pcie(torch, dev, sync=barrier)
    pw = N // 32 * WIRE_BITS
    # pack the operands on the device (one launch of the library's own kernel is not available for a bare repack, so
    # the host-side serialiser of the ABI does it; outside the timed region -- it is the shim's gather step)
    ha_np = fhe.pack_bits(WIRE_BITS, a.cpu().numpy().view(np.uint64))
    hb_np = fhe.pack_bits(WIRE_BITS, b.cpu().numpy().view(np.uint64))
    hap, hbp, hcp = pinned(torch.int32, pw), pinned(torch.int32, pw), pinned(torch.int32, pw)
    hap.copy_(torch.from_numpy(ha_np.view(np.int32)))
    hbp.copy_(torch.from_numpy(hb_np.view(np.int32)))
    del ha_np, hb_np

    def check_packed(hc):
        k = 2048  # unpack a sample on the host and compare with the device-resident result
        got = fhe.unpack_bits(WIRE_BITS, hc[:k].numpy().view(np.uint32))
        return bool((got == c[:k].cpu().numpy().view(np.uint64)).all()) and \
            bool((fhe.unpack_bits(WIRE_BITS, hc[-k:].numpy().view(np.uint32)) == c[-k:].cpu().numpy().view(np.uint64)).all())

    e2e_value, same = e2e_run(hap, hbp, hcp, lambda x, y, z: plan.mul_packed(WIRE_BITS, x, y, out=z), check_packed)
    del hap, hbp, hcp
    e2e32_value, same32 = e2e_run(pinned(torch.int32, N, a), pinned(torch.int32, N, b), pinned(torch.int32, N),
                                  lambda x, y, z: plan.mul_u32(x, y, out=z),
                                  lambda hc: bool(torch.equal(hc.to(dev).to(torch.int64), c)))
    e2e64_value, same64 = e2e_run(pinned(torch.int64, N, a), pinned(torch.int64, N, b), pinned(torch.int64, N),
                                  lambda x, y, z: plan.mul(x, y, out=z), lambda hc: bool(torch.equal(hc.to(dev), c)))
    pcie_all = {k: gather_list(v) for k, v in pcie.items() if k != "mib"}

    boot = measure_bootstrap(fhe, torch, dist, dev, rank, world, local, quick=quick)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_kind = peaks()
    alg_bytes = 3 * N * 8 * batch
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    # the other roofline of the north star ("the slower of the integer-mulmod and HBM rooflines"): modular
    # multiplications per launch (1.5 N log2 N + 2N per polymul) against the Shoup-modmul rate a register-only
    # microbenchmark reaches on this GPU at these clocks (fhe_int_peak, measured now)
    modmul_per_polymul = 3 * (N // 2) * (N.bit_length() - 1) + 2 * N
    modmul_rate = modmul_per_polymul * batch / (kern_ms * 1e-3)
    h2d_p, d2h_p = 2 * pw * 4 * batch, pw * 4 * batch
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
        "traffic": NCU_TRAFFIC_BYTES if batch == BATCH else None, "traffic_source": NCU_TRAFFIC_SOURCE,
        "peak_source": peak_kind, "kernel": "ntt_kernel<Fermat32,10,5,MUL,u64>",
        "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_ms, "kernel_ms_event_pairs": kern_ms_pairs,
        "kernel_ms_source": "timed region / launches in it (CUDA events on the launching stream; one launch per step)",
        "int": {
            "bound": "int32 Shoup modmul (3 IMAD on the fmaheavy pipe)", "achieved": modmul_rate / 1e12, "peak": modmul_peak / 1e12,
            "unit": "T modmul/s", "frac": modmul_rate / modmul_peak, "modmul_per_polymul": modmul_per_polymul,
            "modmul_executed_per_polymul": MODMUL_EXECUTED if N == 1024 else None,
            "executed_frac": (MODMUL_EXECUTED * batch / (kern_ms * 1e-3)) / modmul_peak if N == 1024 else None,
            "peak_source": "fhe_int_peak(1) microbenchmark, this run",
            "note": "frac counts SURVEY 8d's algorithmic modmuls (17408 per polymul) against the Shoup-modmul peak; the radix-4 Fermat32 kernel executes 12288 twiddle products (executed_frac).  ncu (profiles/r2_polymul_n1024_q65537_u64_fermat32_ncu_full_d.csv): fmaheavy (IMAD) pipe 73 % active (radix-2 Small32 kernel: 81 %), ALU pipe 56 %, LSU 24 %, issue slots 64 %, long-scoreboard the top stall, DRAM 1.573 GB per launch for 1.611 GB algorithmic: co-limited by HBM, the integer pipe and issue",
        },
        "u32_device_format": {
            "value": value_u32, "unit": "polymul/s", "algorithmic_bytes_per_polymul": 3 * N * 4,
            "hbm_frac": (3 * N * 4 * value_u32 / world / 1e9) / hbm_peak,
            "modmul_frac": value_u32 / world * modmul_per_polymul / modmul_peak, "binding": "int32 Shoup modmul",
            "matches_u64_result": u32_same, "kernel": "ntt_kernel<Fermat32,10,5,MUL,u32>",
        },
        "sustained": sustained,
        "bootstrap": boot.get("roofline"),
    }
    line = {
        "metric": "NTT polymul/s", "value": value, "unit": "polymul/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 words; 32-bit lazy Shoup/Montgomery arithmetic for q<2^22",
        "data": "synthetic",
        "config": workload_config(batch, world),
        "roofline": roofline,
        "e2e": {
            "value": e2e_value, "unit": "polymul/s", "h2d_bytes_per_step": h2d_p, "d2h_bytes_per_step": d2h_p,
            "steps": e2e_steps, "matches_device_result": same, "host_affinity": numa_note,
            "call": f"fhe_rq_mul_packed (bit-packed wire, {WIRE_BITS} bits per coefficient), pinned host buffers, "
                    "chunked H2D / one kernel launch per chunk / D2H on three streams",
            "h2d_gbs_achieved": h2d_p * e2e_value / world / batch / 1e9,
            "pcie_microbench_gbs_per_rank": pcie_all,
            # what the measured copy rates allow (2 packed operands in, 1 out per polymul): from the H2D-only rates (an upper
            # bound: the D2H stream shares the link) and from the rates with both directions busy
            "pcie_bound_polymul_per_s": {"h2d_alone": sum(pcie_all["h2d_gbs"]) * 1e9 / (2 * pw * 4),
                                         "both_directions_busy": sum(pcie_all["bidir_each_gbs"]) * 1e9 / (2 * pw * 4)},
            "u32_wire": {"value": e2e32_value, "h2d_bytes_per_step": 2 * N * 4 * batch, "d2h_bytes_per_step": N * 4 * batch,
                         "matches_device_result": same32, "call": "fhe_rq_mul_u32"},
            "u64_wire": {"value": e2e64_value, "h2d_bytes_per_step": 2 * N * 8 * batch, "d2h_bytes_per_step": N * 8 * batch,
                         "matches_device_result": same64, "call": "fhe_rq_mul"},
            "bootstrap": boot["e2e"],
        },
        "gpu_launches": int(launches),
        "clocks": clocks,
        "bootstrap": boot,
    }
    if not args.no_cpu:
        base, (xa, xb, xc) = cpu_polymul_baseline(3.0 if quick else 12.0)
        # the timed kernel is also checked against the CPU port on the baseline's sample
        k = min(xa.shape[0], 4096)
        got = plan.mul(np.ascontiguousarray(xa[:k]), np.ascontiguousarray(xb[:k]))
        base["gpu_matches_on_sample"] = bool((got == xc[:k]).all())
        line["cpu_baseline"] = base
    if not args.no_extras and world == 1:
        try:
            import bench_extras

            del a, b, c
            torch.cuda.empty_cache()
            extras = bench_extras.run(fhe, dev, quick=quick, cpu=not args.no_cpu)
            line["extras"] = extras
            # compact copies inside the key the driver keeps
            roofline["sweep"] = bench_extras.compact_sweep(extras)
            for key in ("bfv", "extprod"):
                if key in extras.get("compact", {}):
                    roofline[key] = extras["compact"][key]
        except Exception as ex:  # extras never invalidate the headline line
            line["extras"] = {"error": repr(ex)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
