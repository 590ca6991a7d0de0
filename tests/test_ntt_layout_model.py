"""CPU checks of the NTT pass structure (tools/ntt_model.py, the executable model of csrc/ntt_core.cuh's index
algebra): the register-blocked passes reproduce the reference loops (arith/src/ntt.rs:44-110) for every pass split
the kernels use, and every shared-memory layout of the default shapes is bank-conflict-free under the kernels'
padding -- the property the pass split of ntt_core.cuh (NttShape::g) was chosen for."""
import os
import random
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
import ntt_model as M  # noqa: E402


@pytest.mark.parametrize("logn,loge", [(6, 3), (9, 5), (10, 5), (10, 4), (11, 4), (11, 5), (12, 5), (13, 5), (13, 4), (14, 5)])
def test_pass_structure_matches_reference_loops(logn, loge):
    n = 1 << logn
    roots, roots_inv, n_inv = M.tables(M.Q, n)
    rnd = random.Random(logn * 10 + loge)
    a = [rnd.randrange(M.Q) for _ in range(n)]
    f = M.model_fwd(a, M.Q, roots, logn, loge)
    assert f == M.ref_ntt(a, M.Q, roots)
    assert M.model_fwd(f, M.Q, roots, logn, loge, True, roots_inv, n_inv) == a


def test_split_rule():
    # pass 0 takes the remainder from three passes on (ntt_core.cuh: FRONT), even split up to two
    assert M.split(13, 5) == [3, 5, 5] and M.split(14, 5) == [4, 5, 5] and M.split(12, 5) == [2, 5, 5]
    assert M.split(10, 4) == [2, 4, 4] and M.split(14, 4) == [2, 4, 4, 4]
    assert M.split(10, 5) == [5, 5] and M.split(9, 5) == [5, 4] and M.split(7, 4) == [4, 3]


@pytest.mark.parametrize("logn", [10, 12, 13, 14, 15])
def test_default_32bit_shapes_are_conflict_free(logn):
    # 32 coefficients per thread (the default of every 32-bit degree but N=2048)
    # ... under PadRule (csrc/ntt_kernels.cuh): four pad words per 32, the last pass read and written as aligned
    # 128-bit rows of a thread's 32 consecutive words
    assert M.pad_rule(logn, 5, 4) == (4, True)
    gs, worst = M.conflicts(logn, 5, word_bytes=4)
    assert max(worst.values()) == 1, (gs, worst)


@pytest.mark.parametrize("logn", [10, 11, 12, 13, 14])
def test_default_64bit_shapes_are_conflict_free(logn):
    gs, worst = M.conflicts(logn, 4, word_bytes=8)
    assert max(worst.values()) == 1, (gs, worst)


def test_n2048_16_per_thread_keeps_a_two_way_conflict():
    # documented in DESIGN 4.1: no additive padding serves both the middle and the last pass at 16 per thread
    gs, worst = M.conflicts(11, 4, word_bytes=4)
    assert gs == [3, 4, 4] and worst == {0: 1, 1: 2, 2: 1}
