"""CPU replay of the CUDA kernels' per-thread code (TEST INFRASTRUCTURE ONLY; see emu_ntt.cpp)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libemu.so")
_SRCS = ["emu_ntt.cpp"]


def build() -> str:
    deps = [os.path.join(_HERE, s) for s in _SRCS]
    csrc = os.path.join(_HERE, "..", "..", "fhe_study_b200", "csrc")
    deps += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".hpp"))]
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-x", "c++"]
                              + [os.path.join(_HERE, s) for s in _SRCS] + ["-o", _SO])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        U, P, I = C.c_uint64, C.c_void_p, C.c_int
        L.emu_ntt.argtypes = [I, U, U, I, I, P, P, P, P, I]
        L.emu_ntt.restype = I
        L.emu_xp_digit.argtypes = [U, U, P, P]
        L.emu_xp_digit.restype = I
        L.emu_plan.argtypes = [U, U, P, P, P, P]
        L.emu_plan.restype = I
        L.emu_modmul.argtypes = [I, U, U, U]
        L.emu_modmul.restype = U
        _lib = L
    return _lib
