#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 ring-arithmetic hot path (BASELINE.json configs[1]).

A "step" is one pass of the hot path over one batch of synthetic input: `batch` independent negacyclic
R_q polymuls c = intt(ntt(a) . ntt(b)) at N=1024, q=65537 (the reference's NTT prime; ring_nq::mul,
arith/src/ring_nq.rs:586-607).  `value` is whole-job polymul/s with inputs resident in HBM; `e2e` is the
same metric through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region); `roofline`
is the polymul kernel against the measured HBM peak; `cpu_baseline` is the oracle port timed on this
box's host cores; `extras` carries the rest of the sweep and the TFHE / BFV paths.

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
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of this exact workload (ncu --set full capture)
NCU_TRAFFIC_BYTES = 1073804000 + 499474432
NCU_TRAFFIC_SOURCE = "profiles/r1_polymul_n1024_q65537_ncu_full_e.csv"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


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
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def cpu_polymul_baseline(target_s: float = 12.0):
    """Oracle port (oracle/fhe_oracle.c: u128 % q butterflies, arith/src/ntt.rs + ring_nq.rs) on all host cores."""
    import numpy as np

    import oracle

    cores = os.cpu_count() or 1
    a = oracle.uniform(1, (cores * 8, N), Q)
    b = oracle.uniform(2, (cores * 8, N), Q)
    t0 = time.perf_counter()
    oracle.rq_mul_batch(Q, N, a, b, threads=cores)
    dt = time.perf_counter() - t0
    rate = a.shape[0] / dt
    sample = int(max(cores * 8, min(rate * target_s, 4_000_000)))
    a = oracle.uniform(3, (sample, N), Q)
    b = oracle.uniform(4, (sample, N), Q)
    t0 = time.perf_counter()
    c = oracle.rq_mul_batch(Q, N, a, b, threads=cores)
    dt = time.perf_counter() - t0
    return {
        "value": sample / dt, "unit": "polymul/s", "cores": cores, "kind": "port",
        "sample": f"{sample} polymuls N={N} q={Q} (oracle C port, OpenMP over the batch), {dt:.2f} s",
    }, (a, b, c)


def run_reference(args, emit):
    """--impl reference: the reference's CPU algorithm (oracle port; no Rust toolchain exists here) on all host
    cores, same metric/config.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    cores = os.cpu_count() or 1
    per_step = cores * 1024  # ~0.1 s of CPU work per step
    a = oracle.uniform(3, (per_step, N), Q)
    b = oracle.uniform(4, (per_step, N), Q)
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
        "config": {
            "workload": f"BASELINE configs[1]: batched Rq negacyclic NTT polymul, N={N}, q={Q}, batch {BATCH} per GPU",
            "n": N, "q": Q, "batch_per_step_sampled": per_step,
            "note": "CPU arm: each step is a bounded sample of the workload (the full 65536-polymul batch takes ~0.4 s "
                    "per step on these cores); throughput is per polymul, so the sample size does not change the metric",
        },
        "cpu_baseline": {"value": v, "unit": "polymul/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} polymuls per step, oracle C port of arith/src/ntt.rs + ring_nq.rs"},
        "e2e": {"value": v, "unit": "polymul/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def measure_bootstrap(fhe, torch, dist, dev, rank, world, quick):
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
    for _ in range(2):
        fhe.bootstrap(n, k, K, table, cts, kn, out=out)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fhe.bootstrap(n, k, K, table, cts, kn, out=out)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # end to end through fhe_bootstrap with pinned HOST buffers (chunked H2D / compute / D2H overlap inside the call)
    hct = torch.empty(cts.shape, dtype=torch.int64).pin_memory()
    hout = torch.empty(cts.shape, dtype=torch.int64).pin_memory()
    hct.copy_(cts)
    fhe.bootstrap(n, k, K, table, hct, kn, out=hout)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fhe.bootstrap(n, k, K, table, hct, kn, out=hout)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    e2e_ok = bool(torch.equal(hout.to(dev), out))
    return {
        "e2e": {"value": world * batch * steps / (ms_e2e * 1e-3), "unit": "bootstraps/s",
                "h2d_bytes_per_step": batch * (kn + 1) * 8, "d2h_bytes_per_step": batch * (kn + 1) * 8,
                "matches_device_result": e2e_ok},
        "metric": "TFHE bootstraps/s (as executed: mod_switch + rotate + sample_extract + key_switch)",
        "value": world * batch * steps / (ms * 1e-3), "unit": "bootstraps/s", "n_gpus": world,
        "batch_per_gpu": batch, "steps": steps, "ms_per_step": ms / steps, "scaling": "weak",
        "params": {"n": n, "k": k, "l": l, "ksk_bytes": kn * l * (kn + 1) * 8},
        "key_broadcast_ms": bcast_ms,
        "u64_mac_per_s": world * batch * steps * kn * l * (kn + 1) / (ms * 1e-3),
    }


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
    plan = fhe.NttPlan(Q, N)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    a = torch.randint(0, Q, (batch, N), dtype=torch.int64, device=dev, generator=g)
    b = torch.randint(0, Q, (batch, N), dtype=torch.int64, device=dev, generator=g)
    c = torch.empty_like(a)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") + per-launch kernel time ("roofline") ------------------
    for _ in range(args.warmup):
        plan.mul(a, b, out=c)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = fhe.launch_count()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for s, e in ev:
        s.record()
        plan.mul(a, b, out=c)
        e.record()
    t_end.record()
    barrier()
    launches = fhe.launch_count() - l0
    total_ms = t_start.elapsed_time(t_end)
    kern_ms = statistics.mean(s.elapsed_time(e) for s, e in ev)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * batch * args.steps / (total_ms_max * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers ---------------------------------------
    # Two wire formats of the same call: u64 words (fhe_rq_mul) and packed u32 words (fhe_rq_mul_u32, q <= 2^32:
    # half the PCIe bytes; the Rust shim gathers Vec<Zq>.v into either layout at the same cost).  The step is
    # PCIe-bound in both, so the packed one is the end-to-end headline and the u64 one is reported beside it.
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_run(host_dtype, call):
        ha = torch.empty((batch, N), dtype=host_dtype).pin_memory()
        hb = torch.empty((batch, N), dtype=host_dtype).pin_memory()
        hc = torch.empty((batch, N), dtype=host_dtype).pin_memory()
        ha.copy_(a)
        hb.copy_(b)
        call(ha, hb, hc)  # warm the staging pool
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e2e_steps):
            call(ha, hb, hc)  # H2D(a,b) -> kernel -> D2H(c); returns when c is on the host
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = bool(torch.equal(hc.to(dev).to(torch.int64), c))
        return world * batch * e2e_steps / (float(t.item()) * 1e-3), ok

    e2e64_value, same64 = e2e_run(torch.int64, lambda x, y, z: plan.mul(x, y, out=z))
    e2e_value, same = e2e_run(torch.int32, lambda x, y, z: plan.mul_u32(x, y, out=z))
    clocks = sampler.stop()  # sampled across the timed regions (device-resident steps and end-to-end steps)

    boot = measure_bootstrap(fhe, torch, dist, dev, rank, world, quick=args.steps < 20)

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
    modmul_peak = fhe.int_peak(1)
    modmul_rate = modmul_per_polymul * batch / (kern_ms * 1e-3)
    line = {
        "metric": "NTT polymul/s", "value": value, "unit": "polymul/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64 words; 32-bit lazy Shoup/Montgomery arithmetic for q<2^22",
        "data": "synthetic",
        "config": {
            "workload": f"BASELINE configs[1]: batched Rq negacyclic NTT polymul, N={N}, q={Q}, batch {batch} per GPU",
            "batch_per_gpu": batch, "n": N, "q": Q,
            "l2_policy": f"inputs+outputs {alg_bytes / 2**20:.0f} MiB per step, larger than the 126 MB L2",
            "parallelism": f"independent polynomials sharded over {world} GPU(s), no collective",
            "host_affinity": numa_note,
        },
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": NCU_TRAFFIC_BYTES if batch == BATCH else None, "traffic_source": NCU_TRAFFIC_SOURCE,
            "peak_source": peak_kind, "kernel": "ntt_kernel<Small32,10,5,MUL>",
            "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_ms,
        },
        "roofline_int": {
            "bound": "int32 Shoup modmul (3 IMAD)", "achieved": modmul_rate / 1e12, "peak": modmul_peak / 1e12,
            "unit": "T modmul/s", "frac": modmul_rate / modmul_peak, "modmul_per_polymul": modmul_per_polymul,
            "peak_source": "fhe_int_peak microbenchmark, this run",
        },
        "e2e": {
            "value": e2e_value, "unit": "polymul/s", "h2d_bytes_per_step": 2 * N * 4 * batch,
            "d2h_bytes_per_step": N * 4 * batch, "steps": e2e_steps, "matches_device_result": same,
            "call": "fhe_rq_mul_u32 (packed 32-bit wire, q <= 2^32), pinned host buffers",
        },
        "e2e_u64_wire": {
            "value": e2e64_value, "unit": "polymul/s", "h2d_bytes_per_step": 2 * N * 8 * batch,
            "d2h_bytes_per_step": N * 8 * batch, "steps": e2e_steps, "matches_device_result": same64,
            "call": "fhe_rq_mul (u64 words), pinned host buffers",
        },
        "gpu_launches": int(launches),
        "clocks": clocks,
        "bootstrap": boot,
    }
    if not args.no_cpu:
        base, (xa, xb, xc) = cpu_polymul_baseline()
        # the timed kernel is also checked against the CPU port on the baseline's sample
        k = min(xa.shape[0], 4096)
        got = plan.mul(np.ascontiguousarray(xa[:k]), np.ascontiguousarray(xb[:k]))
        base["gpu_matches_on_sample"] = bool((got == xc[:k]).all())
        line["cpu_baseline"] = base
    if not args.no_extras and world == 1:
        try:
            import bench_extras

            line["extras"] = bench_extras.run(fhe, dev, quick=args.steps < 20, cpu=not args.no_cpu)
        except Exception as ex:  # extras never invalidate the headline line
            line["extras"] = {"error": repr(ex)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
