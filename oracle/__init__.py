"""CPU oracle loader -- TEST INFRASTRUCTURE ONLY.

Thin ctypes/numpy binding over ``oracle/fhe_oracle.c`` (a plain-C restatement of the reference's
ring-arithmetic hot path; every C function cites the reference file:line it follows).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  Nothing under ``fhe_study_b200/`` imports it: the product path
fails loudly when the CUDA library is missing instead of falling back to this code.

Parity pinning: see the header of fhe_oracle.c and tests/test_oracle_kats.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfhe_oracle.so")

U64 = C.c_uint64
P = C.c_void_p


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, no GPU needed). Returns the path of the shared object."""
    src = os.path.join(_HERE, "fhe_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


def _declare(L):
    def d(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    I = C.c_int
    U32 = C.c_uint32
    D = C.c_double
    d("orc_zq_from_u64", U64, U64, U64)
    d("orc_zq_from_f64", U64, U64, D)
    d("orc_zq_add", U64, U64, U64, U64)
    d("orc_zq_sub", U64, U64, U64, U64)
    d("orc_zq_neg", U64, U64, U64)
    d("orc_zq_mul", U64, U64, U64, U64)
    d("orc_zq_exp", U64, U64, U64, U64)
    d("orc_zq_mod_switch", U64, U64, U64, U64)
    d("orc_zq_decompose", None, U64, U64, U32, U32, P)
    d("orc_exp_mod", U64, U64, U64, U64)
    d("orc_inv_mod", U64, U64, U64)
    d("orc_primitive_root_of_unity", U64, U64, U64)
    d("orc_ntt_tables", U64, U64, U64, P, P)
    d("orc_ntt", None, U64, U64, P, P)
    d("orc_intt", None, U64, U64, P, P)
    d("orc_rq_fold", U64, U64, U64, P, U64)
    d("orc_rq_from_vec_u64", U64, U64, U64, P, U64, P)
    d("orc_rq_from_vec_f64", U64, U64, U64, P, U64, P)
    d("orc_rq_from_vec_i64", U64, U64, U64, P, U64, P)
    d("orc_rq_mul", None, U64, U64, P, P, P, I, I, P)
    d("orc_rq_mul_batch", None, U64, U64, P, P, P, U64, I)
    d("orc_ntt_batch", None, U64, U64, P, P, U64, I, I)
    d("orc_rq_addsub", None, U64, U64, P, P, P, I)
    d("orc_rq_mul_u64", None, U64, U64, P, U64, P)
    d("orc_rq_remodule", None, U64, P, U64, P)
    d("orc_rq_mod_switch", None, U64, U64, P, U64, P)
    d("orc_rq_mul_div_round", None, U64, U64, P, U64, U64, P)
    d("orc_rq_decompose", None, U64, U64, P, U32, U32, P)
    d("orc_t64_mod_switch", U64, U64, U64)
    d("orc_t64_mul_div_round", U64, U64, U64, U64)
    d("orc_tn_mul", None, U64, P, P, P)
    d("orc_tn_mul_fast", None, U64, P, P, P)
    d("orc_tn_mul_batch", None, U64, P, P, P, U64, I)
    d("orc_tn_left_rotate", None, U64, P, U64, P)
    d("orc_tn_decompose", None, U64, P, U32, P)
    d("orc_tn_addsub", None, U64, P, P, P, I)
    d("orc_tn_mul_u64", None, U64, P, U64, P)
    d("orc_tn_mod_switch", None, U64, P, U64, P)
    d("orc_tn_mul_div_round", None, U64, P, U64, U64, P)
    d("orc_r_naive_mul", None, U64, P, P, P)
    d("orc_r_fold", U64, U64, P, U64)
    d("orc_r_mul_div_round", None, U64, U64, P, U64, U64, U64, P)
    d("orc_r_mul_to_rq", None, U64, P, P, U64, P)
    d("orc_bfv_tensor", None, U64, U64, U64, P, P, P, P, P)
    d("orc_bfv_relinearize_204", None, U64, U64, U64, P, P, P, P, P)
    d("orc_bfv_mul", None, U64, U64, U64, U64, P, P, P, P)
    d("orc_bfv_mul_batch", None, U64, U64, U64, U64, P, P, P, P, U64, I)
    d("orc_tglwe_mul_tn", None, U64, U64, P, P, P)
    d("orc_tggsw_extprod", None, U64, U64, P, P, P)
    d("orc_tggsw_extprod_fast", None, U64, U64, P, P, P)
    d("orc_tggsw_cmux", None, U64, U64, P, P, P, P)
    d("orc_tggsw_cmux_fast", None, U64, U64, P, P, P, P)
    d("orc_extprod_batch", None, U64, U64, P, P, P, U64, I)
    d("orc_tglwe_left_rotate", None, U64, U64, P, U64, P)
    d("orc_tglwe_sample_extraction", None, U64, U64, P, U64, P)
    d("orc_tlwe_key_switch", None, U64, U64, U32, P, P, P)
    d("orc_tlwe_mod_switch", None, U64, P, U64, P)
    d("orc_compute_lookup_table", None, U64, U64, U64, P)
    d("orc_blind_rotation_as_executed", None, U64, U64, P, U64, P, P)
    d("orc_blind_rotation_as_written", None, U64, U64, P, U64, P, P, P)
    d("orc_bootstrapping", None, U64, U64, P, P, P, U64, P)
    d("orc_bootstrapping_batch", None, U64, U64, P, P, P, U64, P, U64, I)
    d("orc_key_switch_batch", None, U64, U64, U32, P, P, P, U64, I)
    d("orc_fill_uniform_u64", None, U64, P, U64, U64)
    d("orc_tglwe_keygen", None, U64, U64, U64, P)
    d("orc_tglwe_encrypt_s", None, U64, U64, U64, D, P, P, I, P)
    d("orc_tglwe_decrypt", None, U64, U64, P, P, P)
    d("orc_tglwe_encode", None, U64, U64, P, P)
    d("orc_tglwe_decode", None, U64, U64, P, P)
    d("orc_tggsw_encrypt_s", None, U64, U64, U64, D, P, P, I, P)
    d("orc_tlwe_keygen", None, U64, U64, P)
    d("orc_tlwe_encrypt_s", None, U64, U64, D, P, U64, I, P)
    d("orc_tlwe_decrypt", U64, U64, P, P)
    d("orc_tlwe_new_ksk", None, U64, U64, U64, U32, D, P, P, I, P)
    d("orc_cmux_chain", None, U64, U64, U64, P, P, P, I, P)
    d("orc_bootstrap_chain", None, U64, U64, U64, P, P, P, P, U64, I, P)
    d("orc_tlwe_new_ksk_ctr", None, U64, U64, U64, U32, D, P, P, I, P)
    d("orc_tlwe_encrypt_ctr", None, U64, U64, D, P, P, U64, I, P)
    d("orc_bfv_encrypt_ctr", None, U64, U64, U64, U64, D, P, P, U64, P)
    d("orc_tglwe_encrypt_ctr", None, U64, U64, U64, D, P, P, U64, I, P)
    d("orc_tggsw_encrypt_s_ctr", None, U64, U64, U64, D, P, P, I, P)
    d("orc_glev_rq_mul", None, U64, U64, U64, U64, P, P, P)
    d("orc_glwe_rq_key_switch", None, U64, U64, U64, U32, U32, P, P, P)
    d("orc_glwe_rq_mod_switch", None, U64, U64, U64, P, U64, P)
    d("orc_glwe_rq_keygen", None, U64, U64, U64, U64, P)
    d("orc_glwe_rq_encrypt_s", None, U64, U64, U64, U64, D, P, P, P)
    d("orc_glwe_rq_decrypt", None, U64, U64, U64, P, P, P)
    d("orc_glwe_rq_new_ksk", None, U64, U64, U64, U64, U32, U32, D, P, P, P)
    d("orc_bfv_keygen", None, U64, U64, U64, P, P)
    d("orc_bfv_encrypt", None, U64, U64, U64, U64, P, P, P)
    d("orc_bfv_decrypt", None, U64, U64, U64, P, P, P)
    d("orc_bfv_rlk_key", None, U64, U64, U64, U64, P, P)
    d("orc_bfv_keygen_ctr", None, U64, U64, U64, D, P, P)
    d("orc_bfv_rlk_key_ctr", None, U64, U64, U64, U64, D, P, P)
    d("orc_bfv_mul_const", None, U64, U64, U64, U64, P, P, P, P)
    d("orc_ckks_keygen_ctr", None, U64, U64, U64, D, P, P)
    d("orc_ckks_encrypt_ctr", None, U64, U64, U64, D, P, P, U64, P)
    d("orc_ckks_decrypt", None, U64, U64, P, P, U64, P)
    d("orc_ckks_addsub", None, U64, U64, P, P, U64, I, P)


# ---------------------------------------------------------------------------------------------
# numpy conveniences
# ---------------------------------------------------------------------------------------------
def u64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


def i64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def uniform(seed: int, shape, modulus: int = 0) -> np.ndarray:
    """SplitMix64 stream (the synthetic-input generator SURVEY 8d names), optionally reduced mod `modulus`."""
    out = np.empty(int(np.prod(shape)), dtype=np.uint64)
    lib().orc_fill_uniform_u64(seed, ptr(out), out.size, modulus)
    return out.reshape(shape)


def ntt_tables(q: int, n: int):
    roots = np.empty(n, dtype=np.uint64)
    roots_inv = np.empty(n, dtype=np.uint64)
    n_inv = lib().orc_ntt_tables(q, n, ptr(roots), ptr(roots_inv))
    return roots, roots_inv, int(n_inv)


def ntt(q: int, n: int, a, inverse: bool = False, threads: int = 1) -> np.ndarray:
    a = u64(a)
    out = np.empty_like(a)
    lib().orc_ntt_batch(q, n, ptr(a), ptr(out), a.size // n, int(inverse), threads)
    return out


def rq_mul(q: int, n: int, a, b, a_is_evals=False, b_is_evals=False, want_evals=False):
    a, b = u64(a), u64(b)
    c = np.empty(n, dtype=np.uint64)
    ev = np.empty(n, dtype=np.uint64)
    lib().orc_rq_mul(q, n, ptr(a), ptr(b), ptr(c), int(a_is_evals), int(b_is_evals), ptr(ev))
    return (c, ev) if want_evals else c


def rq_mul_batch(q: int, n: int, a, b, threads: int = 1) -> np.ndarray:
    a, b = u64(a), u64(b)
    c = np.empty_like(a)
    lib().orc_rq_mul_batch(q, n, ptr(a), ptr(b), ptr(c), a.size // n, threads)
    return c


def rq_from_vec_u64(q: int, n: int, v) -> np.ndarray:
    v = u64(v)
    out = np.empty(max(v.size, 1), dtype=np.uint64)
    ln = lib().orc_rq_from_vec_u64(q, n, ptr(v), v.size, ptr(out))
    return out[:ln].copy()


def rq_addsub(q: int, n: int, a, b, op: int) -> np.ndarray:
    a = u64(a)
    b = u64(b) if b is not None else a
    c = np.empty_like(a)
    af, bf, cf = a.reshape(-1), b.reshape(-1), c.reshape(-1)
    for i in range(a.size // n):
        lib().orc_rq_addsub(q, n, ptr(af[i * n:]), ptr(bf[i * n:]), ptr(cf[i * n:]), op)
    return c


def tn_mul(n: int, a, b, threads: int = 1) -> np.ndarray:
    a, b = u64(a), u64(b)
    c = np.empty_like(a)
    lib().orc_tn_mul_batch(n, ptr(a), ptr(b), ptr(c), a.size // n, threads)
    return c


def tn_left_rotate(n: int, a, h: int) -> np.ndarray:
    a = u64(a)
    c = np.empty_like(a)
    lib().orc_tn_left_rotate(n, ptr(a), h, ptr(c))
    return c


def extprod(n: int, k: int, tggsw, ct, fast: bool = True) -> np.ndarray:
    tggsw, ct = u64(tggsw), u64(ct)
    glwe = (k + 1) * n
    out = np.empty_like(ct)
    cf, of = ct.reshape(-1), out.reshape(-1)  # any batch shape: rows are addressed on the flat views
    f = lib().orc_tggsw_extprod_fast if fast else lib().orc_tggsw_extprod
    for i in range(ct.size // glwe):
        f(n, k, ptr(tggsw), ptr(cf[i * glwe:]), ptr(of[i * glwe:]))
    return out


def cmux(n: int, k: int, tggsw, ct1, ct2, fast: bool = True) -> np.ndarray:
    tggsw, ct1, ct2 = u64(tggsw), u64(ct1), u64(ct2)
    glwe = (k + 1) * n
    out = np.empty_like(ct1)
    f1, f2, of = ct1.reshape(-1), ct2.reshape(-1), out.reshape(-1)
    f = lib().orc_tggsw_cmux_fast if fast else lib().orc_tggsw_cmux
    for i in range(ct1.size // glwe):
        f(n, k, ptr(tggsw), ptr(f1[i * glwe:]), ptr(f2[i * glwe:]), ptr(of[i * glwe:]))
    return out


def key_switch(kn_in: int, kn_out: int, l: int, ksk, ct, threads: int = 1) -> np.ndarray:
    ksk, ct = u64(ksk), u64(ct)
    batch = ct.size // (kn_in + 1)
    out = np.empty(batch * (kn_out + 1), dtype=np.uint64)
    lib().orc_key_switch_batch(kn_in, kn_out, l, ptr(ksk), ptr(ct), ptr(out), batch, threads)
    return out


def bootstrapping(n: int, k: int, ksk, table, c, c_kn: int, threads: int = 1) -> np.ndarray:
    ksk, table, c = u64(ksk), u64(table), u64(c)
    batch = c.size // (c_kn + 1)
    out = np.empty(batch * (k * n + 1), dtype=np.uint64)
    lib().orc_bootstrapping_batch(n, k, ptr(ksk), ptr(table), ptr(c), c_kn, ptr(out), batch, threads)
    return out


def cmux_chain(n: int, k: int, bsk, acc, h, negacyclic: bool = False) -> np.ndarray:
    """acc_b <- cmux(bsk[j], acc_b, rotate(acc_b, h[b][j])) for j < steps; bsk = steps flat TGGSWs, h = [batch][steps]."""
    bsk, acc, h = u64(bsk), u64(acc), u64(h)
    glwe = (k + 1) * n
    batch = acc.size // glwe
    steps = h.size // batch
    out = np.empty_like(acc)
    for i in range(batch):
        lib().orc_cmux_chain(n, k, steps, ptr(bsk), ptr(acc.reshape(-1)[i * glwe:]), ptr(h.reshape(-1)[i * steps:]),
                             int(negacyclic), ptr(out.reshape(-1)[i * glwe:]))
    return out


def bootstrap_chain(n: int, k: int, steps: int, bsk, ksk, table, c, c_kn: int, mode: int) -> np.ndarray:
    bsk, table, c = u64(bsk), u64(table), u64(c)
    ksk = u64(ksk) if ksk is not None else None
    kn = k * n
    batch = c.size // (c_kn + 1)
    out = np.empty((batch, kn + 1), dtype=np.uint64)
    for i in range(batch):
        lib().orc_bootstrap_chain(n, k, steps, ptr(bsk), ptr(ksk) if ksk is not None else None, ptr(table),
                                  ptr(c.reshape(-1)[i * (c_kn + 1):]), c_kn, mode, ptr(out[i]))
    return out


def glwe_rq_key_switch(q: int, n: int, k: int, beta: int, l: int, ksk, ct) -> np.ndarray:
    """GLWE<Rq>::key_switch (gfhe/src/glwe.rs:126-137) for a batch of GLWEs of (k+1)*n words."""
    ksk, ct = u64(ksk), u64(ct)
    glwe = (k + 1) * n
    out = np.empty_like(ct)
    for i in range(ct.size // glwe):
        lib().orc_glwe_rq_key_switch(q, n, k, beta, l, ptr(ksk), ptr(ct.reshape(-1)[i * glwe:]), ptr(out.reshape(-1)[i * glwe:]))
    return out


def glev_rq_mul(q: int, n: int, k: int, l: int, glev, v) -> np.ndarray:
    """impl Mul<Vec<R>> for GLev<R> (gfhe/src/glev.rs:67-80), R = Rq; v = batch x l polys."""
    glev, v = u64(glev), u64(v)
    batch = v.size // (l * n)
    out = np.empty((batch, (k + 1) * n), dtype=np.uint64)
    for i in range(batch):
        lib().orc_glev_rq_mul(q, n, k, l, ptr(glev), ptr(v.reshape(-1)[i * l * n:]), ptr(out[i]))
    return out


def tlwe_new_ksk_ctr(seed: int, kn_in: int, kn_out: int, l: int, sigma: float, sk, new_sk, uniform_mask: bool = True) -> np.ndarray:
    """KSK with the counter-based sampler the device generator reproduces (orc_tlwe_new_ksk_ctr)."""
    sk, new_sk = u64(sk), u64(new_sk)
    out = np.empty(kn_in * l * (kn_out + 1), dtype=np.uint64)
    lib().orc_tlwe_new_ksk_ctr(seed, kn_in, kn_out, l, float(sigma), ptr(sk), ptr(new_sk), int(uniform_mask), ptr(out))
    return out


def tggsw_encrypt_s_ctr(seed: int, n: int, k: int, sigma: float, sk, m, uniform_mask: bool = True) -> np.ndarray:
    """TGGSW::encrypt_s with the counter-based sampler the device generator reproduces."""
    sk, m = u64(sk), u64(m)
    out = np.empty((k + 1) * 64 * (k + 1) * n, dtype=np.uint64)
    lib().orc_tggsw_encrypt_s_ctr(seed, n, k, float(sigma), ptr(sk), ptr(m), int(uniform_mask), ptr(out))
    return out


def lookup_table(n: int, k: int, t: int) -> np.ndarray:
    out = np.empty((k + 1) * n, dtype=np.uint64)
    lib().orc_compute_lookup_table(n, k, t, ptr(out))
    return out


def bfv_mul(q: int, n: int, t: int, pq: int, rlk, a, b, threads: int = 1) -> np.ndarray:
    rlk, a, b = u64(rlk), u64(a), u64(b)
    out = np.empty_like(a)
    lib().orc_bfv_mul_batch(q, n, t, pq, ptr(rlk), ptr(a), ptr(b), ptr(out), a.size // (2 * n), threads)
    return out


def bfv_keygen_ctr(seed: int, q: int, n: int, sigma: float):
    """BFV::new_key (bfv/src/lib.rs:120-140) with the counter-based sampler the device reproduces: (sk, pk[2n])."""
    sk, pk = np.empty(n, dtype=np.uint64), np.empty(2 * n, dtype=np.uint64)
    lib().orc_bfv_keygen_ctr(seed, q, n, float(sigma), ptr(sk), ptr(pk))
    return sk, pk


def bfv_rlk_key_ctr(seed: int, q: int, n: int, p: int, sigma: float, sk) -> np.ndarray:
    """BFV::rlk_key (bfv/src/lib.rs:202-225) with the counter-based sampler: rlk[2n] mod p*q."""
    sk = u64(sk)
    out = np.empty(2 * n, dtype=np.uint64)
    lib().orc_bfv_rlk_key_ctr(seed, q, n, p, float(sigma), ptr(sk), ptr(out))
    return out


def bfv_mul_const(q: int, n: int, t: int, pq: int, rlk, c, m) -> np.ndarray:
    """BFV::mul_const (bfv/src/lib.rs:189-200) for a batch: c = batch x 2n, m = batch x n (mod t)."""
    rlk, c, m = u64(rlk), u64(c), u64(m)
    out = np.empty_like(c)
    for i in range(c.size // (2 * n)):
        lib().orc_bfv_mul_const(q, n, t, pq, ptr(rlk), ptr(c.reshape(-1)[i * 2 * n:]), ptr(m.reshape(-1)[i * n:]),
                                ptr(out.reshape(-1)[i * 2 * n:]))
    return out


def ckks_keygen_ctr(seed: int, q: int, n: int, sigma: float):
    sk, pk = np.empty(n, dtype=np.uint64), np.empty(2 * n, dtype=np.uint64)
    lib().orc_ckks_keygen_ctr(seed, q, n, float(sigma), ptr(sk), ptr(pk))
    return sk, pk


def ckks_encrypt_ctr(seed: int, q: int, n: int, sigma: float, pk, msgs) -> np.ndarray:
    """CKKS::encrypt (ckks/src/lib.rs:66-84); msgs = batch x n int64 (elements of R)."""
    pk = u64(pk)
    msgs = np.ascontiguousarray(msgs, dtype=np.int64)
    batch = msgs.size // n
    out = np.empty((batch, 2 * n), dtype=np.uint64)
    lib().orc_ckks_encrypt_ctr(seed, q, n, float(sigma), ptr(pk), ptr(msgs), batch, ptr(out))
    return out


def ckks_decrypt(q: int, n: int, sk, ct) -> np.ndarray:
    """CKKS::decrypt (ckks/src/lib.rs:86-94): batch x n int64 (centred representatives)."""
    sk, ct = u64(sk), u64(ct)
    batch = ct.size // (2 * n)
    out = np.empty((batch, n), dtype=np.int64)
    lib().orc_ckks_decrypt(q, n, ptr(sk), ptr(ct), batch, ptr(out))
    return out


def ckks_addsub(q: int, n: int, c0, c1, sub: bool) -> np.ndarray:
    c0, c1 = u64(c0), u64(c1)
    out = np.empty_like(c0)
    lib().orc_ckks_addsub(q, n, ptr(c0), ptr(c1), c0.size // (2 * n), int(sub), ptr(out))
    return out
