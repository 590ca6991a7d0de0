"""Secondary measurements printed under `extras` in bench.py's JSON line (device-resident, CUDA events):
the NTT / INTT / polymul sweep of BASELINE configs[1] and -- once built -- the TFHE and BFV paths of
configs[2..4].  Each entry carries its own roofline fraction against the measured HBM peak."""
from __future__ import annotations

import json
import os

ROOT = os.path.dirname(os.path.abspath(__file__))
Q17 = 65537
Q62 = 0x3FFFFFFFFFFF0001


def _hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def _time(fn, reps, warm=3):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps  # ms


def ntt_sweep(fhe, dev, quick):
    import torch

    peak = _hbm_peak()
    out = []
    reps = 3 if quick else 20
    for q in (Q17, Q62):
        for logn in (10, 11, 12, 13, 14):
            n = 1 << logn
            batch = (1 << 26) // n  # 512 MiB per u64 operand
            plan = fhe.NttPlan(q, n)
            for fmt, dt, wb in (("u64", torch.int64, 8),) + ((("u32", torch.int32, 4),) if q <= 2**32 else ()):
                a = torch.randint(0, min(q, 2**31 - 1 if wb == 4 else 2**62), (batch, n), dtype=dt, device=dev)
                b = torch.randint(0, min(q, 2**31 - 1 if wb == 4 else 2**62), (batch, n), dtype=dt, device=dev)
                c = torch.empty_like(a)
                row = {"q": q, "n": n, "batch": batch, "format": fmt, "bytes_per_coeff": wb}
                fns = (("ntt", lambda: plan.ntt(a, out=c), 2), ("intt", lambda: plan.intt(a, out=c), 2),
                       ("polymul", lambda: plan.mul(a, b, out=c), 3)) if wb == 8 else \
                      (("ntt", lambda: plan.ntt_u32(a, out=c), 2), ("intt", lambda: plan.intt_u32(a, out=c), 2),
                       ("polymul", lambda: plan.mul_u32(a, b, out=c), 3))
                for name, fn, nbuf in fns:
                    ms = _time(fn, reps)
                    gbs = nbuf * n * wb * batch / (ms * 1e-3) / 1e9
                    row[name] = {"per_s": batch / (ms * 1e-3), "ms": ms, "hbm_gbs": gbs, "hbm_frac": gbs / peak}
                out.append(row)
                del a, b, c
    # batch sweep at N=1024 (BASELINE configs[1]: batch 1..64k), device-resident; small batches are launch-latency bound
    plan = fhe.NttPlan(Q17, 1024)
    batch_sweep = []
    for batch in (1, 16, 256, 4096, 65536):
        a = torch.randint(0, Q17, (batch, 1024), dtype=torch.int64, device=dev)
        b = torch.randint(0, Q17, (batch, 1024), dtype=torch.int64, device=dev)
        c = torch.empty_like(a)
        ms = _time(lambda: plan.mul(a, b, out=c), 20 if batch >= 4096 else 100, warm=5)
        batch_sweep.append({"batch": batch, "us_per_call": ms * 1e3, "polymul_per_s": batch / (ms * 1e-3),
                            "note": "working set fits the 126 MB L2" if batch * 3 * 8192 < 126e6 else "HBM-resident"})
        del a, b, c
    # batch-1 latency (BASELINE configs[0]: the crate's own test path), device-resident
    plan = fhe.NttPlan(Q17, 1024)
    a = torch.randint(0, Q17, (1, 1024), dtype=torch.int64, device=dev)
    c = torch.empty_like(a)
    lat = _time(lambda: plan.mul(a, a, out=c), 200, warm=20)
    # integer rooflines: modmuls per second of each sweep point against the register-only Shoup-modmul peak
    peaks = {"imad32_per_s": fhe.int_peak(0), "shoup32_modmul_per_s": fhe.int_peak(1), "shoup64_modmul_per_s": fhe.int_peak(2)}
    for row in out:
        n, logn = row["n"], row["n"].bit_length() - 1
        pk = peaks["shoup32_modmul_per_s"] if row["q"] < 2**30 else peaks["shoup64_modmul_per_s"]
        work = {"ntt": n // 2 * logn, "intt": n // 2 * logn + n, "polymul": 3 * (n // 2) * logn + 2 * n}
        for name in ("ntt", "intt", "polymul"):
            row[name]["modmul_frac"] = row[name]["per_s"] * work[name] / pk
            row[name]["roofline_frac"] = max(row[name]["hbm_frac"], row[name]["modmul_frac"])
    return {"sweep": out, "batch_sweep_n1024": batch_sweep, "polymul_n1024_batch1_us": lat * 1e3, "int_peaks": peaks}




# ---------------------------------------------------------------------------------------------------------
# TFHE / BFV paths (BASELINE configs[2..4]); device-resident buffers, CUDA events, each with a bounded CPU
# sample of the oracle port on all host cores beside it.
# ---------------------------------------------------------------------------------------------------------
def _u64_rand(torch, shape, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    return torch.randint(-(2**63), 2**63 - 1, shape, dtype=torch.int64, device=dev, generator=g)


def tfhe_paths(fhe, dev, quick, cpu=True):
    import time

    import numpy as np
    import torch

    import oracle

    cores = os.cpu_count() or 1
    res = {}
    reps = 2 if quick else 5
    # --- config 3/4: TGGSW x TGLWE external product and CMux, reference parameter sets -----------------
    for name, n, k, batch in (("P4a_n64_k4", 64, 4, 4096), ("P4b_n1024_k1", 1024, 1, 1024 if quick else 4096)):
        glwe = (k + 1) * n
        tggsw = _u64_rand(torch, ((k + 1) * 64 * glwe,), dev, 1)
        g = fhe.Tggsw(n, k, tggsw)
        ct1 = _u64_rand(torch, (batch, glwe), dev, 2)
        ct2 = _u64_rand(torch, (batch, glwe), dev, 3)
        out = torch.empty_like(ct1)
        ms_e = _time(lambda: g.extprod(ct1, out=out), reps, warm=1)
        ms_c = _time(lambda: g.cmux(ct1, ct2, out=out), reps, warm=1)
        row = {"n": n, "k": k, "batch": batch, "extprod_per_s": batch / (ms_e * 1e-3), "cmux_per_s": batch / (ms_c * 1e-3),
               "extprod_ms": ms_e, "cmux_ms": ms_c}
        # integer roofline (limb-NTT formulation): 2 primes x [(k+1)*64 forward + 2(k+1) inverse
        # transforms] + 2 primes x (k+1)*64 digits x 2(k+1) units x n MACs, against the measured Shoup-modmul peak
        logn = n.bit_length() - 1
        modmul = 2 * ((k + 1) * 64 * (n // 2) * logn + 2 * (k + 1) * ((n // 2) * logn + n)) + 2 * (k + 1) * 64 * 2 * (k + 1) * n
        row["modmul_per_unit"] = modmul
        pk = fhe.int_peak(1)
        row["int_roofline_frac"] = row["extprod_per_s"] * modmul / pk
        # the same against SURVEY 8d's count (ONE transform set: (k+1)*64 forward + 2(k+1) inverse + 2(k+1)^2*64*n MACs);
        # the kernel runs the set under two CRT primes, which is what `modmul` above counts
        survey = (k + 1) * 64 * (n // 2) * logn + 2 * (k + 1) * ((n // 2) * logn + n) + 2 * (k + 1) * (k + 1) * 64 * n
        row["modmul_per_unit_survey_8d"] = survey
        row["int_roofline_frac_survey_8d_count"] = row["extprod_per_s"] * survey / pk
        # what the fused kernel EXECUTES since r2: stages 0-2 of every forward digit transform are table lookups
        # (extprod_fused.cu), so a forward transform multiplies in logn - 3 stages only
        executed = 2 * ((k + 1) * 64 * (n // 2) * (logn - 3) + 2 * (k + 1) * ((n // 2) * logn + n)) + 2 * (k + 1) * 64 * 2 * (k + 1) * n
        row["modmul_executed_per_unit"] = executed
        row["int_roofline_frac_executed"] = row["extprod_per_s"] * executed / pk
        row["hbm_frac"] = row["extprod_per_s"] * 2 * (k + 1) * n * 8 / (_hbm_peak() * 1e9)
        if n == 1024 and batch > 1024:
            # a batch that does not fill whole waves of accumulator pairs: the library falls back to one accumulator per
            # 256-thread CTA there (extprod_fused.cu: xp_use_pair)
            ms_s = _time(lambda: g.extprod(ct1[:1024], out=out[:1024]), reps, warm=1)
            row["batch_1024"] = {"extprod_per_s": 1024 / (ms_s * 1e-3), "extprod_ms": ms_s}
        if cpu:
            sample = max(1, min(batch, cores * (4 if n <= 64 else 1)))
            hg = tggsw.cpu().numpy().view(np.uint64)
            hc = ct1[:sample].cpu().numpy().view(np.uint64).reshape(-1).copy()
            ho = np.empty_like(hc)
            t0 = time.perf_counter()
            oracle.lib().orc_extprod_batch(n, k, oracle.ptr(hg), oracle.ptr(hc), oracle.ptr(ho), sample, cores)
            dt = time.perf_counter() - t0
            same = bool((g.extprod(ct1[:sample].contiguous()).cpu().numpy().view(np.uint64).reshape(-1) == ho).all())
            row["cpu_extprod_per_s"] = sample / dt
            row["cpu_sample"] = f"{sample} external products, O(N^2) u128 schoolbook port, {cores} threads, {dt:.2f} s"
            row["gpu_matches_cpu_sample"] = same
        res[name] = row
        del g, tggsw, ct1, ct2, out
    # --- config 5: bootstrapping as executed (n=1024, k=1, l=64; 537 MB KSK resident) ---------------------
    n, k, kn, l = 1024, 1, 1024, 64
    batch = 2048 if quick else 8192
    ksk = _u64_rand(torch, (kn * l * (kn + 1),), dev, 4)
    K = fhe.Ksk(kn, kn, l, ksk)
    table = _u64_rand(torch, ((k + 1) * n,), dev, 5)  # dense (non-trivial) table: dense digits in the key switch
    cts = _u64_rand(torch, (batch, kn + 1), dev, 6)
    out = torch.empty_like(cts)
    ms_b = _time(lambda: fhe.bootstrap(n, k, K, table, cts, kn, out=out), reps, warm=1)
    ms_k = _time(lambda: K.key_switch(cts, out=out), reps, warm=1)
    row = {"n": n, "k": k, "l": l, "batch": batch, "bootstraps_per_s": batch / (ms_b * 1e-3), "bootstrap_ms": ms_b,
           "key_switch_per_s": batch / (ms_k * 1e-3), "ksk_bytes": int(ksk.numel() * 8),
           "u64_mac_per_s": batch * kn * l * (kn + 1) / (ms_b * 1e-3)}
    # tensor roofline: every u64 MAC is 8 byte-plane int8 MACs on the tensor cores (ks_tc.cu) = 16 int8 ops
    row["int8_pops"] = row["u64_mac_per_s"] * 16 / 1e15
    row["tensor_roofline_frac_vs_4.5_pops_nominal"] = row["int8_pops"] / 4.5
    if cpu:
        sample = cores * 2
        hk = ksk.cpu().numpy().view(np.uint64)
        ht = table.cpu().numpy().view(np.uint64)
        hc = cts[:sample].cpu().numpy().view(np.uint64).reshape(-1).copy()
        t0 = time.perf_counter()
        ho = oracle.bootstrapping(n, k, hk, ht, hc, kn, threads=cores)
        dt = time.perf_counter() - t0
        got = fhe.bootstrap(n, k, K, table, cts[:sample].contiguous(), kn).cpu().numpy().view(np.uint64).reshape(-1)
        row["cpu_bootstraps_per_s"] = sample / dt
        row["cpu_sample"] = f"{sample} bootstraps (as executed), oracle port, {cores} threads, {dt:.2f} s"
        row["gpu_matches_cpu_sample"] = bool((got == ho).all())
    res["P5_bootstrap_as_executed"] = row
    # --- the same 537 MB key sampled on the device instead of uploaded ----------------------
    sk_bits = (_u64_rand(torch, (2, kn), dev, 9) & 1)
    t0 = time.perf_counter()
    Kg = fhe.Ksk.generate(kn, kn, l, sk_bits[0].contiguous(), sk_bits[1].contiguous(), sigma=3.2, seed=5)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    gen = {"seconds_incl_tensor_core_relayout": gen_s, "ksk_bytes": int(kn * l * (kn + 1) * 8)}
    if cpu:
        h0 = sk_bits.cpu().numpy().view(np.uint64)
        t0 = time.perf_counter()
        ref = oracle.tlwe_new_ksk_ctr(5, kn, kn, l, 3.2, h0[0].copy(), h0[1].copy(), True)
        gen["cpu_seconds"] = time.perf_counter() - t0
        gen["cpu_note"] = f"same sampler, oracle port, OpenMP over the rows on {cores} threads"
        gen["gpu_matches_cpu"] = bool((Kg.export() == ref).all())
        del ref
    res["P5b_ksk_generated_on_device"] = gen
    del Kg
    # --- SURVEY 8f rank 1 (extension; no reference execution): blind rotation as a CMux chain with one TGGSW per
    #     mask element, accumulator resident on chip for the whole chain, then sample extraction + key switch ----
    steps = 32 if quick else 256
    cb = 296 if quick else 592
    rows = _u64_rand(torch, ((k + 1) * 64 * (k + 1) * n,), dev, 7)
    bsk = [fhe.Tggsw(n, k, rows) for _ in range(steps)]  # distinct resident copies (12 MB each), as a real BSK
    c2 = _u64_rand(torch, (cb, steps + 1), dev, 8)
    o2 = torch.empty((cb, kn + 1), dtype=torch.int64, device=dev)
    ms_p = _time(lambda: fhe.bootstrap_chain(n, k, bsk, table, c2, steps, mode=1, ksk=K, out=o2), 2, warm=1)
    res["P6_bootstrap_cmux_chain_extension"] = {
        "n": n, "k": k, "steps": steps, "batch": cb, "ms": ms_p, "cmux_per_s": cb * steps / (ms_p * 1e-3),
        "bootstraps_per_s_at_these_steps": cb / (ms_p * 1e-3),
        "note": "mod_switch(2n) + X^-b table + `steps` on-chip CMuxes (l=64, beta=2 as tggsw.rs:49-50) + sample_extract + "
                "key_switch; the reference never executes such a chain (SURVEY F3)"}
    del K, ksk, cts, out, bsk
    return res


def bfv_path(fhe, dev, quick, cpu=True):
    import time

    import numpy as np
    import torch

    import oracle

    cores = os.cpu_count() or 1
    q, n, t = Q17, 16, 2
    pq = q * q * q
    res = {}
    for batch in (4096, 1 << 20):
        a = torch.randint(0, q, (batch, 2 * n), dtype=torch.int64, device=dev)
        b = torch.randint(0, q, (batch, 2 * n), dtype=torch.int64, device=dev)
        rlk = torch.randint(0, pq, (2 * n,), dtype=torch.int64, device=dev)
        out = torch.empty_like(a)
        ms = _time(lambda: fhe.bfv_mul_relin(q, n, t, pq, rlk, a, b, out=out), 5 if quick else 20)
        row = {"q": q, "n": n, "t": t, "batch": batch, "mul_relin_per_s": batch / (ms * 1e-3), "ms": ms,
               "hbm_gbs": 48 * n * batch / (ms * 1e-3) / 1e9}
        if batch == 4096:
            # one call of 4096 ciphertexts is a single ~6 us kernel behind ~10 us of host call path: replaying 16 such
            # calls from one CUDA graph shows the rate the device sustains at this batch size
            try:
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    fhe.use_torch_stream()
                    fhe.bfv_mul_relin(q, n, t, pq, rlk, a, b, out=out)
                    with torch.cuda.graph(g, stream=side):
                        fhe.use_torch_stream()
                        for _ in range(16):
                            fhe.bfv_mul_relin(q, n, t, pq, rlk, a, b, out=out)
                torch.cuda.current_stream().wait_stream(side)
                fhe.use_torch_stream()
                gms = _time(g.replay, 20)
                row["graph_replay_16_calls"] = {"mul_relin_per_s": 16 * batch / (gms * 1e-3), "us_per_call": gms * 1e3 / 16}
            except Exception as ex:
                fhe.use_torch_stream()
                row["graph_replay_16_calls"] = {"error": repr(ex)[:200]}
        if cpu and batch == 4096:
            ha, hb, hr = (x.cpu().numpy().view(np.uint64) for x in (a, b, rlk))
            reps = 50
            t0 = time.perf_counter()
            for _ in range(reps):
                ho = oracle.bfv_mul(q, n, t, pq, hr, ha.reshape(-1), hb.reshape(-1), threads=cores)
            dt = time.perf_counter() - t0
            row["cpu_mul_relin_per_s"] = reps * batch / dt
            row["cpu_sample"] = f"{reps} x {batch} ct muls, oracle port, {cores} threads, {dt:.2f} s"
            row["gpu_matches_cpu_sample"] = bool((out.cpu().numpy().view(np.uint64).reshape(-1) == ho).all())
        res[f"batch_{batch}"] = row
    return res


def tn_mul_path(fhe, dev, quick):
    import torch

    res = {}
    for n, batch in ((1024, 8192), (64, 65536)):
        a = _u64_rand(torch, (batch, n), dev, 7)
        b = _u64_rand(torch, (batch, n), dev, 8)
        c = torch.empty_like(a)
        ms = _time(lambda: fhe.tn_mul(n, a, b, out=c), 3 if quick else 10)
        res[f"n{n}"] = {"batch": batch, "tn_mul_per_s": batch / (ms * 1e-3), "ms": ms}
    return res


def gfhe_path(fhe, dev, quick, cpu=True):
    """GLWE<Rq>::key_switch at the reference's test parameters (gfhe/src/glwe.rs:582-594:
    q=65537, n=128, k=16, beta=2, l=16), batched, with a CPU sample of the oracle port beside it."""
    import time

    import numpy as np
    import torch

    import oracle

    q, n, k, beta, l = Q17, 128, 16, 2, 16
    batch = 256 if quick else 2048
    glwe = (k + 1) * n
    g = torch.Generator(device=dev).manual_seed(21)
    ksk = torch.randint(0, q, (k * l * glwe,), dtype=torch.int64, device=dev, generator=g)
    ct = torch.randint(0, q, (batch, glwe), dtype=torch.int64, device=dev, generator=g)
    K = fhe.RqGlev(fhe.NttPlan(q, n), k, k * l, ksk)
    out = torch.empty_like(ct)
    ms = _time(lambda: K.key_switch(beta, l, ct, out=out), 2 if quick else 5, warm=1)
    row = {"q": q, "n": n, "k": k, "beta": beta, "l": l, "batch": batch, "key_switch_per_s": batch / (ms * 1e-3), "ms": ms}
    if cpu:
        sample = min(batch, 4 * (os.cpu_count() or 1))
        hk = ksk.cpu().numpy().view(np.uint64)
        hc = ct[:sample].cpu().numpy().view(np.uint64).copy()
        t0 = time.perf_counter()
        want = oracle.glwe_rq_key_switch(q, n, k, beta, l, hk, hc)
        dt = time.perf_counter() - t0
        row["cpu_key_switch_per_s"] = sample / dt
        row["cpu_sample"] = f"{sample} key switches, oracle port (single thread: the reference is sequential), {dt:.2f} s"
        row["gpu_matches_cpu_sample"] = bool((out[:sample].cpu().numpy().view(np.uint64) == want).all())
    return row


def compact_sweep(extras):
    """The NTT sweep as short rows for the `roofline` key of the bench line: [q bits, n, format, M polymul/s,
    fraction of HBM peak, fraction of the measured Shoup-modmul peak, fraction of the binding (slower) roofline],
    plus the same binding fraction for the forward and inverse transforms alone."""
    rows = []
    for r in extras.get("ntt", {}).get("sweep", []):
        pm = r["polymul"]
        rows.append({"q_bits": int(r["q"]).bit_length(), "n": r["n"], "fmt": r["format"], "M_polymul_per_s": round(pm["per_s"] / 1e6, 3),
                     "hbm": round(pm["hbm_frac"], 3), "modmul": round(pm["modmul_frac"], 3), "frac": round(pm["roofline_frac"], 3),
                     "ntt_frac": round(r["ntt"]["roofline_frac"], 3), "intt_frac": round(r["intt"]["roofline_frac"], 3)})
    return {"rows": rows, "frac_is": "max(hbm, modmul): achieved / the slower of the HBM and integer-modmul rooflines",
            "min_polymul_frac": min((x["frac"] for x in rows), default=None),
            "peaks": extras.get("ntt", {}).get("int_peaks")}


def run(fhe, dev, quick=False, cpu=True):
    res = {}
    for name, fn in (("ntt", lambda: ntt_sweep(fhe, dev, quick)), ("tfhe", lambda: tfhe_paths(fhe, dev, quick, cpu)),
                     ("bfv", lambda: bfv_path(fhe, dev, quick, cpu)), ("tn_mul", lambda: tn_mul_path(fhe, dev, quick)),
                     ("gfhe", lambda: gfhe_path(fhe, dev, quick, cpu))):
        try:
            res[name] = fn()
        except Exception as ex:  # one failing extra must not hide the others
            res[name] = {"error": repr(ex)}
    compact = {}
    try:
        t = res["tfhe"]
        compact["extprod"] = {k: {"extprod_per_s": t[k]["extprod_per_s"], "cmux_per_s": t[k]["cmux_per_s"], "batch": t[k]["batch"],
                                  "int_roofline_frac": t[k]["int_roofline_frac"], "int_roofline_frac_executed": t[k]["int_roofline_frac_executed"],
                                  "int_roofline_frac_survey_8d_count": t[k]["int_roofline_frac_survey_8d_count"], "hbm_frac": t[k]["hbm_frac"],
                                  "cpu_extprod_per_s": t[k].get("cpu_extprod_per_s")} for k in ("P4a_n64_k4", "P4b_n1024_k1")}
    except Exception:
        pass
    try:
        compact["extprod"]["cmux_chain_n1024_k1"] = {kk: t["P6_bootstrap_cmux_chain_extension"][kk] for kk in ("steps", "batch", "cmux_per_s")}
    except Exception:
        pass
    try:
        compact["bfv"] = res["bfv"]
    except Exception:
        pass
    res["compact"] = compact
    return res
