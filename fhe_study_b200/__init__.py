"""fhe_study_b200 -- B200 (sm_100a) implementation of the ring-arithmetic hot path of arnaucube/fhe-study.

The product is ``libfhe_b200.so`` (hand-written CUDA behind the C ABI of ``include/fhe_b200.h``).  This
package is the thin Python host layer over that ABI used by the tests and ``bench.py``: batched calls on
numpy arrays (host buffers) or torch CUDA tensors (HBM-resident buffers).  Importing it without the
built library raises; there is no CPU path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import FheError, check, lib, ptr  # noqa: F401

A_IS_EVALS = 1
B_IS_EVALS = 2
B_BROADCAST = 4


def _is_torch(x) -> bool:
    return hasattr(x, "data_ptr")


def _empty_like(x):
    if _is_torch(x):
        import torch

        return torch.empty_like(x)
    return np.empty_like(x)


def _numel(x) -> int:
    return x.numel() if _is_torch(x) else x.size


def _check_u64(*xs):
    for x in xs:
        if x is None:
            continue
        if _is_torch(x):
            import torch

            if x.dtype not in (torch.int64, torch.uint64) or not x.is_contiguous():
                raise TypeError("torch buffers must be contiguous int64/uint64 (bit patterns of u64)")
        else:
            if x.dtype != np.uint64 or not x.flags["C_CONTIGUOUS"]:
                raise TypeError("numpy buffers must be C-contiguous uint64")


def _check_u32(*xs):
    for x in xs:
        if x is None:
            continue
        if _is_torch(x):
            import torch

            if x.dtype not in (torch.int32, torch.uint32) or not x.is_contiguous():
                raise TypeError("torch buffers must be contiguous int32/uint32 (bit patterns of u32)")
        else:
            if x.dtype != np.uint32 or not x.flags["C_CONTIGUOUS"]:
                raise TypeError("numpy buffers must be C-contiguous uint32")


def set_device(device: int) -> None:
    check(lib.fhe_set_device(int(device)))


def use_torch_stream() -> None:
    """Run this thread's library calls on torch's current CUDA stream (so torch.cuda.Event sees them)."""
    import torch

    check(lib.fhe_set_device(torch.cuda.current_device()))
    check(lib.fhe_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)))


def synchronize() -> None:
    check(lib.fhe_synchronize())


def rq_check_canonical(q: int, words) -> int | None:
    """Validation pass for Rq buffers (precondition of every Rq entry point: 0 <= v < q): index of the first
    word >= q, or None."""
    bad = C.c_uint64()
    check(lib.fhe_rq_check_canonical(int(q), ptr(words), _numel(words), C.byref(bad)))
    return None if bad.value == 2**64 - 1 else int(bad.value)


def int_peak(kind: int) -> float:
    """Measured integer-pipe peak (ops/s): 0 = IMAD32, 1 = Shoup modmul 32-bit, 2 = Shoup modmul 64-bit."""
    v = C.c_double()
    check(lib.fhe_int_peak(int(kind), C.byref(v)))
    return float(v.value)


def launch_count() -> int:
    return int(lib.fhe_launch_count())


def pack_bits(bits: int, a: np.ndarray) -> np.ndarray:
    """Host-side serialiser of the bit-packed format: u64 coefficients (last axis a multiple of 32) -> u32 words."""
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.shape[-1] % 32:
        raise ValueError("the last axis must be a multiple of 32 coefficients")
    out = np.empty(a.shape[:-1] + (a.shape[-1] // 32 * int(bits),), dtype=np.uint32)
    check(lib.fhe_pack_bits(int(bits), ptr(a), ptr(out), a.size))
    return out


def unpack_bits(bits: int, w: np.ndarray) -> np.ndarray:
    """Inverse of pack_bits: u32 words (last axis a multiple of `bits`) -> u64 coefficients."""
    w = np.ascontiguousarray(w, dtype=np.uint32)
    if w.shape[-1] % int(bits):
        raise ValueError("the last axis must be a multiple of `bits` words")
    out = np.empty(w.shape[:-1] + (w.shape[-1] // int(bits) * 32,), dtype=np.uint64)
    check(lib.fhe_unpack_bits(int(bits), ptr(w), ptr(out), out.size))
    return out


class NttPlan:
    """(q, n) plan: mirrors the reference's global ``(q,n) -> (roots, roots_inv, n_inv)`` cache
    (arith/src/ntt.rs:18-38).  Raises FheError where the reference panics."""

    def __init__(self, q: int, n: int):
        self.q, self.n = int(q), int(n)
        h = C.c_void_p()
        check(lib.fhe_ntt_plan_create(self.q, self.n, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib.fhe_ntt_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def config(self):
        """Diagnostics: (policy, log2 coefficients per thread, dual, gpark, staged) the plan selected."""
        arr = (C.c_int * 5)()
        check(lib.fhe_ntt_plan_config(self._h, arr))
        return {"policy": arr[0], "loge": arr[1], "dual": arr[2], "gpark": arr[3], "staged": arr[4]}

    def info(self):
        psi, n_inv = C.c_uint64(), C.c_uint64()
        roots = np.empty(self.n, dtype=np.uint64)
        roots_inv = np.empty(self.n, dtype=np.uint64)
        check(lib.fhe_ntt_plan_info(self._h, C.byref(psi), C.byref(n_inv), ptr(roots), ptr(roots_inv)))
        return int(psi.value), int(n_inv.value), roots, roots_inv

    def _batch(self, a) -> int:
        ne = _numel(a)
        if ne % self.n:
            raise ValueError("buffer length is not a multiple of n")
        return ne // self.n

    def ntt(self, a, out=None):
        """NTT::ntt (arith/src/ntt.rs:44-73) on every polynomial of `a`."""
        out = _empty_like(a) if out is None else out
        _check_u64(a, out)
        check(lib.fhe_ntt_fwd(self._h, ptr(a), ptr(out), self._batch(a)))
        return out

    def intt(self, a, out=None):
        """NTT::intt (arith/src/ntt.rs:78-110)."""
        out = _empty_like(a) if out is None else out
        _check_u64(a, out)
        check(lib.fhe_ntt_inv(self._h, ptr(a), ptr(out), self._batch(a)))
        return out

    def mul(self, a, b, out=None, flags: int = 0, evals_out=None):
        """ring_nq::mul (arith/src/ring_nq.rs:586-607); `evals_out` receives the product's cached evals."""
        out = _empty_like(a) if out is None else out
        _check_u64(a, b, out, evals_out)
        if _numel(b) != (self.n if int(flags) & B_BROADCAST else _numel(a)):
            raise ValueError("operand sizes differ")
        check(lib.fhe_rq_mul(self._h, ptr(a), ptr(b), ptr(out), self._batch(a), int(flags), ptr(evals_out)))
        return out


    # packed 32-bit wire format (q <= 2^32): same values, half the bytes over PCIe
    def ntt_u32(self, a, out=None):
        out = _empty_like(a) if out is None else out
        _check_u32(a, out)
        check(lib.fhe_ntt_fwd_u32(self._h, ptr(a), ptr(out), self._batch(a)))
        return out

    def intt_u32(self, a, out=None):
        out = _empty_like(a) if out is None else out
        _check_u32(a, out)
        check(lib.fhe_ntt_inv_u32(self._h, ptr(a), ptr(out), self._batch(a)))
        return out

    def mul_u32(self, a, b, out=None, flags: int = 0, evals_out=None):
        out = _empty_like(a) if out is None else out
        _check_u32(a, b, out, evals_out)
        if _numel(a) != _numel(b):
            raise ValueError("operand sizes differ")
        check(lib.fhe_rq_mul_u32(self._h, ptr(a), ptr(b), ptr(out), self._batch(a), int(flags), ptr(evals_out)))
        return out

    # bit-packed wire format (q < 2^30, n >= 1024): n*bits/32 u32 words per polynomial, read and written by the kernels
    def _batch_packed(self, a, bits) -> int:
        row = self.n // 32 * int(bits)
        ne = _numel(a)
        if ne % row:
            raise ValueError("buffer length is not a multiple of n*bits/32 words")
        return ne // row

    def ntt_packed(self, bits, a, out=None):
        out = _empty_like(a) if out is None else out
        _check_u32(a, out)
        check(lib.fhe_ntt_fwd_packed(self._h, int(bits), ptr(a), ptr(out), self._batch_packed(a, bits)))
        return out

    def intt_packed(self, bits, a, out=None):
        out = _empty_like(a) if out is None else out
        _check_u32(a, out)
        check(lib.fhe_ntt_inv_packed(self._h, int(bits), ptr(a), ptr(out), self._batch_packed(a, bits)))
        return out

    def mul_packed(self, bits, a, b, out=None, flags: int = 0, evals_out=None):
        out = _empty_like(a) if out is None else out
        _check_u32(a, b, out, evals_out)
        check(lib.fhe_rq_mul_packed(self._h, int(bits), ptr(a), ptr(b), ptr(out), self._batch_packed(a, bits), int(flags),
                                    ptr(evals_out)))
        return out


# -------------------------------------------------------------------------------------------------------
# torus / TFHE / BFV / coefficient-wise entry points (thin wrappers; names follow the reference)
# -------------------------------------------------------------------------------------------------------
def _new(like, shape):
    if _is_torch(like):
        import torch

        return torch.empty(shape, dtype=like.dtype, device=like.device)
    return np.empty(shape, dtype=np.uint64)


def tn_mul(n, a, b, out=None):
    """impl Mul<Tn> for Tn (arith/src/ring_torus.rs:251-298): exact negacyclic product mod 2^64."""
    out = _empty_like(a) if out is None else out
    _check_u64(a, b, out)
    check(lib.fhe_tn_mul(int(n), ptr(a), ptr(b), ptr(out), _numel(a) // int(n)))
    return out


def tn_add(a, b, out=None):
    out = _empty_like(a) if out is None else out
    _check_u64(a, b, out)
    check(lib.fhe_tn_add(ptr(a), ptr(b), ptr(out), _numel(a)))
    return out


def tn_sub(a, b, out=None):
    out = _empty_like(a) if out is None else out
    _check_u64(a, b, out)
    check(lib.fhe_tn_sub(ptr(a), ptr(b), ptr(out), _numel(a)))
    return out


def tn_neg(a, out=None):
    out = _empty_like(a) if out is None else out
    _check_u64(a, out)
    check(lib.fhe_tn_neg(ptr(a), ptr(out), _numel(a)))
    return out


def tn_left_rotate(n, a, h, group=1, out=None):
    """Tn::left_rotate / TGLWE::left_rotate; `h` holds one amount per group of `group` polynomials."""
    out = _empty_like(a) if out is None else out
    _check_u64(a, h, out)
    check(lib.fhe_tn_left_rotate(int(n), ptr(a), ptr(h), int(group), ptr(out), _numel(a) // int(n)))
    return out


class Tggsw:
    """Device-resident TGGSW (tfhe/src/tggsw.rs:14), transformed once at load."""

    def __init__(self, n, k, rows):
        self.n, self.k = int(n), int(k)
        _check_u64(rows)
        if _numel(rows) != (self.k + 1) * 64 * (self.k + 1) * self.n:
            raise ValueError("TGGSW must hold (k+1)*64*(k+1)*n words")
        h = C.c_void_p()
        check(lib.fhe_tggsw_load(self.n, self.k, ptr(rows), C.byref(h)))
        self._h = h

    @classmethod
    def generate(cls, n, k, sk, m, sigma=3.2, seed=0, uniform_mask=True, rows_out=None):
        """TGGSW::encrypt_s (tfhe/src/tggsw.rs:17-33) sampled on the device; rows_out optionally receives the rows."""
        self = cls.__new__(cls)
        self.n, self.k = int(n), int(k)
        _check_u64(sk, m, rows_out)
        if _numel(sk) != self.k * self.n or _numel(m) != self.n:
            raise ValueError("sk must hold k*n words and m n words")
        h = C.c_void_p()
        check(lib.fhe_tggsw_generate(self.n, self.k, ptr(sk), ptr(m), float(sigma), int(seed), int(bool(uniform_mask)),
                                     ptr(rows_out), C.byref(h)))
        self._h = h
        return self

    def close(self):
        if getattr(self, "_h", None):
            lib.fhe_tggsw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _batch(self, ct):
        return _numel(ct) // ((self.k + 1) * self.n)

    def extprod(self, ct, out=None):
        """impl Mul<TGLWE> for TGGSW (tggsw.rs:45-62)."""
        out = _empty_like(ct) if out is None else out
        _check_u64(ct, out)
        check(lib.fhe_extprod(self._h, ptr(ct), ptr(out), self._batch(ct)))
        return out

    def cmux(self, ct1, ct2, out=None):
        """TGGSW::cmux (tggsw.rs:39-41)."""
        out = _empty_like(ct1) if out is None else out
        _check_u64(ct1, ct2, out)
        check(lib.fhe_cmux(self._h, ptr(ct1), ptr(ct2), ptr(out), self._batch(ct1)))
        return out


class Ksk:
    """Device-resident key-switching key (tfhe/src/tlwe.rs:84-100)."""

    def __init__(self, kn_in, kn_out, l, rows):
        self.kn_in, self.kn_out, self.l = int(kn_in), int(kn_out), int(l)
        _check_u64(rows)
        if _numel(rows) != self.kn_in * self.l * (self.kn_out + 1):
            raise ValueError("KSK must hold kn_in*l*(kn_out+1) words")
        h = C.c_void_p()
        check(lib.fhe_ksk_load(self.kn_in, self.kn_out, self.l, ptr(rows), C.byref(h)))
        self._h = h

    @classmethod
    def generate(cls, kn_in, kn_out, l, sk, new_sk, sigma=3.2, seed=0, uniform_mask=True):
        """TLWE::new_ksk (tfhe/src/tlwe.rs:84-100) generated on the device (counter-based sampler, see fhe_b200.h)."""
        self = cls.__new__(cls)
        self.kn_in, self.kn_out, self.l = int(kn_in), int(kn_out), int(l)
        _check_u64(sk, new_sk)
        if _numel(sk) != self.kn_in or _numel(new_sk) != self.kn_out:
            raise ValueError("sk / new_sk must hold kn_in / kn_out words")
        h = C.c_void_p()
        check(lib.fhe_ksk_generate(self.kn_in, self.kn_out, self.l, ptr(sk), ptr(new_sk), float(sigma), int(seed),
                                   int(bool(uniform_mask)), C.byref(h)))
        self._h = h
        return self

    def export(self):
        """The key rows as a numpy array (layout of the constructor's `rows`)."""
        out = np.empty(self.kn_in * self.l * (self.kn_out + 1), dtype=np.uint64)
        check(lib.fhe_ksk_export(self._h, ptr(out)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib.fhe_ksk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def key_switch(self, ct, out=None):
        """TLWE::key_switch (tlwe.rs:101-112)."""
        batch = _numel(ct) // (self.kn_in + 1)
        out = _new(ct, (batch, self.kn_out + 1)) if out is None else out
        _check_u64(ct, out)
        check(lib.fhe_key_switch(self._h, ptr(ct), ptr(out), batch))
        return out


def bootstrap(n, k, ksk: Ksk, table, ct, c_kn, out=None):
    """bootstrapping (tfhe/src/tlwe.rs:150-161) as the reference executes it."""
    batch = _numel(ct) // (int(c_kn) + 1)
    out = _new(ct, (batch, ksk.kn_out + 1)) if out is None else out
    _check_u64(table, ct, out)
    check(lib.fhe_bootstrap(int(n), int(k), ksk._h, ptr(table), ptr(ct), int(c_kn), ptr(out), batch))
    return out


def blind_rotate(n, k, table, ct, c_kn, bsk=None, as_written=False, out=None):
    """blind_rotation (tfhe/src/tlwe.rs:121-148); bsk = list of k Tggsw (needed for as_written and k > 1)."""
    batch = _numel(ct) // (int(c_kn) + 1)
    out = _new(ct, (batch, (int(k) + 1) * int(n))) if out is None else out
    _check_u64(table, ct, out)
    arr = None
    if bsk is not None:
        arr = (C.c_void_p * len(bsk))(*[b._h for b in bsk])
    check(lib.fhe_blind_rotate(int(n), int(k), arr, int(bool(as_written)), ptr(table), ptr(ct), int(c_kn), ptr(out), batch))
    return out


class RqGlev:
    """Device-resident rows of GLWE<Rq> ciphertexts, transformed once at load: a GLev (rows = l, gfhe/src/glev.rs:14)
    or a key-switching key (rows = k*l, gfhe/src/glwe.rs:99-125)."""

    def __init__(self, plan: "NttPlan", k, rows, glwes):
        self.plan, self.k, self.rows, self.n = plan, int(k), int(rows), plan.n
        _check_u64(glwes)
        if _numel(glwes) != self.rows * (self.k + 1) * self.n:
            raise ValueError("expected rows * (k+1) * n words")
        h = C.c_void_p()
        check(lib.fhe_rq_glev_load(plan._h, self.k, self.rows, ptr(glwes), C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib.fhe_rq_glev_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def mul(self, v, out=None):
        """impl Mul<Vec<R>> for GLev<R> (glev.rs:67-80): v = batch x rows polynomials -> batch GLWEs."""
        batch = _numel(v) // (self.rows * self.n)
        out = _new(v, (batch, (self.k + 1) * self.n)) if out is None else out
        _check_u64(v, out)
        check(lib.fhe_rq_glev_mul(self._h, ptr(v), ptr(out), batch))
        return out

    def key_switch(self, beta, l, ct, out=None):
        """GLWE<Rq>::key_switch (glwe.rs:126-137)."""
        out = _empty_like(ct) if out is None else out
        _check_u64(ct, out)
        check(lib.fhe_glwe_rq_key_switch(self._h, int(beta), int(l), ptr(ct), ptr(out), _numel(ct) // ((self.k + 1) * self.n)))
        return out


def _handles(bsk):
    return (C.c_void_p * max(len(bsk), 1))(*[b._h for b in bsk])


def cmux_chain(n, k, bsk, acc, h, negacyclic=False, out=None):
    """acc_b <- cmux(bsk[j], acc_b, X^{-h[b][j]} acc_b) for j < len(bsk): the loop of blind_rotation
    (tfhe/src/tlwe.rs:138-147) for len(bsk) TGGSWs; the accumulator stays on chip for the whole chain.  Extension."""
    batch = _numel(acc) // ((int(k) + 1) * int(n))
    out = _empty_like(acc) if out is None else out
    _check_u64(acc, out) if len(bsk) == 0 else _check_u64(acc, h, out)
    if len(bsk) and _numel(h) != batch * len(bsk):
        raise ValueError("h must hold batch * steps rotation amounts")
    check(lib.fhe_cmux_chain(int(n), int(k), _handles(bsk), len(bsk), int(bool(negacyclic)), ptr(acc),
                             ptr(h) if len(bsk) else None, ptr(out), batch))
    return out


def bootstrap_chain(n, k, bsk, table, ct, c_kn, mode=1, ksk: Ksk = None, out=None):
    """Bootstrapping with one TGGSW per mask element (see fhe_bootstrap_chain in include/fhe_b200.h).  Extension."""
    batch = _numel(ct) // (int(c_kn) + 1)
    width = (ksk.kn_out if ksk is not None else int(k) * int(n)) + 1
    out = _new(ct, (batch, width)) if out is None else out
    _check_u64(table, ct, out)
    check(lib.fhe_bootstrap_chain(int(n), int(k), _handles(bsk), len(bsk), int(mode), ksk._h if ksk is not None else None,
                                  ptr(table), ptr(ct), int(c_kn), ptr(out), batch))
    return out


def tlwe_encrypt(kn, sk, msgs, sigma=3.2, seed=0, uniform_mask=True, out=None):
    """TLWE::encrypt_s (tfhe/src/tlwe.rs:71-74) of already-encoded messages, sampled on the device."""
    batch = _numel(msgs)
    out = _new(msgs, (batch, int(kn) + 1)) if out is None else out
    _check_u64(sk, msgs, out)
    check(lib.fhe_tlwe_encrypt(int(kn), ptr(sk), ptr(msgs), float(sigma), int(seed), int(bool(uniform_mask)), ptr(out), batch))
    return out


def tlwe_decrypt(kn, sk, ct, out=None):
    """TLWE::decrypt (tfhe/src/tlwe.rs:80-82): the phase b - <a, sk> of every ciphertext."""
    batch = _numel(ct) // (int(kn) + 1)
    out = _new(ct, (batch,)) if out is None else out
    _check_u64(sk, ct, out)
    check(lib.fhe_tlwe_decrypt(int(kn), ptr(sk), ptr(ct), ptr(out), batch))
    return out


def tglwe_encrypt(n, k, sk, msgs, sigma=3.2, seed=0, uniform_mask=True, out=None):
    """TGLWE::encrypt_s (tfhe/src/tglwe.rs:76-79) of already-encoded message polynomials, sampled on the device."""
    batch = _numel(msgs) // int(n)
    out = _new(msgs, (batch, (int(k) + 1) * int(n))) if out is None else out
    _check_u64(sk, msgs, out)
    check(lib.fhe_tglwe_encrypt(int(n), int(k), ptr(sk), ptr(msgs), float(sigma), int(seed), int(bool(uniform_mask)), ptr(out), batch))
    return out


def tglwe_decrypt(n, k, sk, ct, out=None):
    """TGLWE::decrypt (tfhe/src/tglwe.rs:86-88): b - sum_i a_i * sk_i."""
    batch = _numel(ct) // ((int(k) + 1) * int(n))
    out = _new(ct, (batch, int(n))) if out is None else out
    _check_u64(sk, ct, out)
    check(lib.fhe_tglwe_decrypt(int(n), int(k), ptr(sk), ptr(ct), ptr(out), batch))
    return out


def torus_decode(p, t):
    """TLWE::decode / TGLWE::decode (tlwe.rs:60-63, tglwe.rs:59-63): round(t * p / u64::MAX) reduced mod t."""
    return rq_remodule(tn_mul_div_round(p, int(t), 2**64 - 1), int(t))


def sample_extract(n, k, ct, h, out=None):
    batch = _numel(ct) // ((int(k) + 1) * int(n))
    out = _new(ct, (batch, int(k) * int(n) + 1)) if out is None else out
    _check_u64(ct, out)
    check(lib.fhe_sample_extract(int(n), int(k), ptr(ct), int(h), ptr(out), batch))
    return out


def tlwe_mod_switch(ct, q2, out=None):
    out = _empty_like(ct) if out is None else out
    _check_u64(ct, out)
    check(lib.fhe_tlwe_mod_switch(ptr(ct), int(q2), ptr(out), _numel(ct)))
    return out


def bfv_tensor(q, n, t, a, b, out=None):
    batch = _numel(a) // (2 * int(n))
    out = _new(a, (batch, 3 * int(n))) if out is None else out
    _check_u64(a, b, out)
    check(lib.fhe_bfv_tensor(int(q), int(n), int(t), ptr(a), ptr(b), ptr(out), batch))
    return out


def bfv_relinearize(q, n, pq, rlk, c012, out=None):
    batch = _numel(c012) // (3 * int(n))
    out = _new(c012, (batch, 2 * int(n))) if out is None else out
    _check_u64(rlk, c012, out)
    check(lib.fhe_bfv_relinearize(int(q), int(n), int(pq), ptr(rlk), ptr(c012), ptr(out), batch))
    return out


def bfv_mul_relin(q, n, t, pq, rlk, a, b, out=None):
    """RLWE::mul (bfv/src/lib.rs:87-90): tensor + relinearize_204."""
    out = _empty_like(a) if out is None else out
    _check_u64(rlk, a, b, out)
    check(lib.fhe_bfv_mul_relin(int(q), int(n), int(t), int(pq), ptr(rlk), ptr(a), ptr(b), ptr(out), _numel(a) // (2 * int(n))))
    return out


def bfv_encrypt(plan: NttPlan, t, pk, m, sigma=3.2, seed=0, out=None):
    """BFV::encrypt (bfv/src/lib.rs:142-160) of a batch of messages mod t, sampled on the device."""
    batch = _numel(m) // plan.n
    out = _new(m, (batch, 2 * plan.n)) if out is None else out
    _check_u64(pk, m, out)
    check(lib.fhe_bfv_encrypt(plan._h, plan.q, plan.n, int(t), ptr(pk), ptr(m), float(sigma), int(seed), ptr(out), batch))
    return out


def bfv_decrypt(plan: NttPlan, t, sk, ct, out=None):
    """BFV::decrypt (bfv/src/lib.rs:164-178) for a batch of RLWEs (2n words each) under one secret key."""
    batch = _numel(ct) // (2 * plan.n)
    out = _new(ct, (batch, plan.n)) if out is None else out
    _check_u64(sk, ct, out)
    check(lib.fhe_bfv_decrypt(plan._h, plan.q, plan.n, int(t), ptr(sk), ptr(ct), ptr(out), batch))
    return out


def bfv_keygen(plan: NttPlan, sigma=3.2, seed=0):
    """BFV::new_key (bfv/src/lib.rs:120-140) sampled on the device: (sk[n], pk[2n]) as numpy arrays."""
    sk, pk = np.empty(plan.n, dtype=np.uint64), np.empty(2 * plan.n, dtype=np.uint64)
    check(lib.fhe_bfv_keygen(plan._h, plan.q, plan.n, float(sigma), int(seed), ptr(sk), ptr(pk)))
    return sk, pk


def bfv_rlk_generate(q, n, p, sk, sigma=3.2, seed=0):
    """BFV::rlk_key (bfv/src/lib.rs:202-225) on the device: rlk[2n] mod p*q."""
    _check_u64(sk)
    out = _new(sk, (2 * int(n),))
    check(lib.fhe_bfv_rlk_generate(int(q), int(n), int(p), float(sigma), int(seed), ptr(sk), ptr(out)))
    return out


def bfv_mul_const(q, n, t, pq, rlk, c, m, out=None):
    """BFV::mul_const (bfv/src/lib.rs:189-200) for a batch of RLWEs and plaintext polynomials mod t."""
    out = _empty_like(c) if out is None else out
    _check_u64(rlk, c, m, out)
    check(lib.fhe_bfv_mul_const(int(q), int(n), int(t), int(pq), ptr(rlk), ptr(c), ptr(m), ptr(out), _numel(c) // (2 * int(n))))
    return out


def ckks_keygen(plan: NttPlan, sigma=3.2, seed=0):
    """CKKS::new_key (ckks/src/lib.rs:46-63) sampled on the device: (sk[n], pk[2n])."""
    sk, pk = np.empty(plan.n, dtype=np.uint64), np.empty(2 * plan.n, dtype=np.uint64)
    check(lib.fhe_ckks_keygen(plan._h, plan.q, plan.n, float(sigma), int(seed), ptr(sk), ptr(pk)))
    return sk, pk


def ckks_encrypt(plan: NttPlan, pk, m, sigma=3.2, seed=0):
    """CKKS::encrypt (ckks/src/lib.rs:66-84); m = batch x n int64 (elements of R); returns batch x 2n u64."""
    m = np.ascontiguousarray(m, dtype=np.int64)
    batch = m.size // plan.n
    out = np.empty((batch, 2 * plan.n), dtype=np.uint64)
    _check_u64(pk)
    check(lib.fhe_ckks_encrypt(plan._h, plan.q, plan.n, ptr(pk), ptr(m), float(sigma), int(seed), ptr(out), batch))
    return out


def ckks_decrypt(plan: NttPlan, sk, ct):
    """CKKS::decrypt (ckks/src/lib.rs:86-94): batch x n int64 (centred representatives mod q)."""
    _check_u64(sk, ct)
    batch = _numel(ct) // (2 * plan.n)
    out = np.empty((batch, plan.n), dtype=np.int64)
    check(lib.fhe_ckks_decrypt(plan._h, plan.q, plan.n, ptr(sk), ptr(ct), ptr(out), batch))
    return out


def ckks_add(q, n, c0, c1, sub=False):
    """CKKS::add / CKKS::sub (ckks/src/lib.rs:113-118; sub as written adds the second components)."""
    out = _empty_like(c0)
    _check_u64(c0, c1, out)
    fn = lib.fhe_ckks_sub if sub else lib.fhe_ckks_add
    check(fn(int(q), int(n), ptr(c0), ptr(c1), ptr(out), _numel(c0) // (2 * int(n))))
    return out


def compute_lookup_table(n, k, t):
    """compute_lookup_table (tfhe/src/tlwe.rs:196-214): the trivial TGLWE of the staircase table, (k+1)*n words."""
    out = np.empty((int(k) + 1) * int(n), dtype=np.uint64)
    check(lib.fhe_compute_lookup_table(int(n), int(k), int(t), ptr(out)))
    return out


def _map1(fn, a, *scalars, pre=()):
    out = _empty_like(a)
    _check_u64(a, out)
    check(fn(*pre, ptr(a), *scalars, ptr(out), _numel(a)))
    return out


def rq_add(q, a, b):
    out = _empty_like(a)
    _check_u64(a, b, out)
    check(lib.fhe_rq_add(int(q), ptr(a), ptr(b), ptr(out), _numel(a)))
    return out


def rq_sub(q, a, b):
    out = _empty_like(a)
    _check_u64(a, b, out)
    check(lib.fhe_rq_sub(int(q), ptr(a), ptr(b), ptr(out), _numel(a)))
    return out


def rq_neg(q, a):
    return _map1(lib.fhe_rq_neg, a, pre=(int(q),))


def rq_mul_u64(q, a, s):
    return _map1(lib.fhe_rq_mul_u64, a, int(s), pre=(int(q),))


def rq_remodule(a, p):
    return _map1(lib.fhe_rq_remodule, a, int(p))


def rq_mod_switch(q, a, p):
    return _map1(lib.fhe_rq_mod_switch, a, int(p), pre=(int(q),))


def rq_mul_div_round(q, a, num, den):
    return _map1(lib.fhe_rq_mul_div_round, a, int(num), int(den), pre=(int(q),))


def rq_from_vec(q, n, v, in_len):
    batch = _numel(v) // int(in_len)
    out = _new(v, (batch, int(n)))
    _check_u64(v, out)
    check(lib.fhe_rq_from_vec(int(q), int(n), ptr(v), int(in_len), ptr(out), batch))
    return out


def rq_decompose(q, n, a, beta, l):
    polys = _numel(a) // int(n)
    out = _new(a, (polys, int(l), int(n)))
    _check_u64(a, out)
    check(lib.fhe_rq_decompose(int(q), int(n), ptr(a), int(beta), int(l), ptr(out), polys))
    return out


def tn_decompose(n, a, l=64):
    polys = _numel(a) // int(n)
    out = _new(a, (polys, int(l), int(n)))
    _check_u64(a, out)
    check(lib.fhe_tn_decompose(int(n), ptr(a), int(l), ptr(out), polys))
    return out


def tn_mod_switch(a, p):
    return _map1(lib.fhe_tn_mod_switch, a, int(p))


def tn_mul_u64(a, s):
    return _map1(lib.fhe_tn_mul_u64, a, int(s))


def tn_mul_div_round(a, num, den):
    return _map1(lib.fhe_tn_mul_div_round, a, int(num), int(den))


# ---- flat on-disk / wire container (include/fhe_b200_file.h) -------------------------------------
FILE_KINDS = {"rq": 1, "tn": 2, "tlwe": 3, "tglwe": 4, "tggsw": 5, "ksk": 6, "rlwe": 7, "rlk": 8, "secret": 9, "glev_rq": 10}
ENC_U64, ENC_U32, ENC_PACKED = 0, 1, 2


def save(path: str, kind: str, words: np.ndarray, words_per_object: int, q: int = 0, n: int = 0, k: int = 0, l: int = 0,
         encoding: int = ENC_U64, bits: int = 0) -> dict:
    """Write `words` (u64 values, count x words_per_object) as one container file.  encoding = ENC_U32 / ENC_PACKED store
    Rq data (q <= 2^32 / q <= 2^bits) in the narrower wire formats of the host-buffer path."""
    from ._capi import FileInfo

    words = np.ascontiguousarray(words, dtype=np.uint64).reshape(-1)
    if words.size % int(words_per_object):
        raise ValueError("payload is not a whole number of objects")
    info = FileInfo(kind=FILE_KINDS[kind], encoding=int(encoding), bits=int(bits), q=int(q), n=int(n), k=int(k), l=int(l),
                    count=words.size // int(words_per_object), words_per_object=int(words_per_object))
    if encoding == ENC_U32:
        payload = words.astype(np.uint32)
        if (payload.astype(np.uint64) != words).any():
            raise ValueError("coefficients do not fit 32 bits")
    elif encoding == ENC_PACKED:
        payload = pack_bits(bits, words.reshape(-1, int(words_per_object))).reshape(-1)
    else:
        payload = words
    check(lib.fhe_file_write(path.encode(), C.byref(info), ptr(payload)))
    return {f[0]: getattr(info, f[0]) for f in FileInfo._fields_}


def load(path: str):
    """Read a container file: (info dict, u64 words shaped count x words_per_object)."""
    from ._capi import FileInfo

    info = FileInfo()
    check(lib.fhe_file_read_info(path.encode(), C.byref(info)))
    raw = np.empty(int(info.payload_bytes), dtype=np.uint8)
    check(lib.fhe_file_read_payload(path.encode(), ptr(raw), raw.size))
    shape = (int(info.count), int(info.words_per_object))
    if info.encoding == ENC_U32:
        words = raw.view(np.uint32).astype(np.uint64).reshape(shape)
    elif info.encoding == ENC_PACKED:
        words = unpack_bits(int(info.bits), raw.view(np.uint32).reshape(shape[0], -1))
    else:
        words = raw.view(np.uint64).reshape(shape).copy()
    return {f[0]: getattr(info, f[0]) for f in FileInfo._fields_}, words
