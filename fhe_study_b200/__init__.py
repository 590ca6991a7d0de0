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


def set_device(device: int) -> None:
    check(lib.fhe_set_device(int(device)))


def use_torch_stream() -> None:
    """Run this thread's library calls on torch's current CUDA stream (so torch.cuda.Event sees them)."""
    import torch

    check(lib.fhe_set_device(torch.cuda.current_device()))
    check(lib.fhe_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream)))


def synchronize() -> None:
    check(lib.fhe_synchronize())


def launch_count() -> int:
    return int(lib.fhe_launch_count())


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
        if _numel(a) != _numel(b):
            raise ValueError("operand sizes differ")
        check(lib.fhe_rq_mul(self._h, ptr(a), ptr(b), ptr(out), self._batch(a), int(flags), ptr(evals_out)))
        return out
