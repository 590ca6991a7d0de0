"""CPU replay of the CUDA NTT kernels' per-thread code (tests/emu) against the oracle.  No GPU needed:
this is what keeps the index / twiddle / lazy-reduction algebra of csrc/ntt_core.cuh honest between
GPU runs.  The real kernels are checked in test_gpu_ntt.py."""
import numpy as np
import pytest

from primes import Q17, Q22, Q30, Q62, Q63


@pytest.fixture(scope="module")
def emu():
    from emu import lib

    return lib()


def _check(emu, orc, kind, q, n, loge):
    a = orc.uniform(n + 1, (n,), q)
    b = orc.uniform(n + 2, (n,), q)
    fa = orc.ntt(q, n, a)
    out = np.zeros(n, dtype=np.uint64)
    assert emu.emu_ntt(kind, q, n, loge, 0, orc.ptr(a), None, orc.ptr(out), None, 0) == 0
    assert (out == fa).all()
    assert emu.emu_ntt(kind, q, n, loge, 1, orc.ptr(fa), None, orc.ptr(out), None, 0) == 0
    assert (out == a).all()
    c, ev = orc.rq_mul(q, n, a, b, want_evals=True)
    ce = np.zeros(n, dtype=np.uint64)
    assert emu.emu_ntt(kind, q, n, loge, 2, orc.ptr(a), orc.ptr(b), orc.ptr(out), orc.ptr(ce), 0) == 0
    assert (out == c).all() and (ce == ev).all()
    fb = orc.ntt(q, n, b)
    for fl, (aa, bb) in {1: (fa, b), 2: (a, fb), 3: (fa, fb)}.items():
        assert emu.emu_ntt(kind, q, n, loge, 2, orc.ptr(aa), orc.ptr(bb), orc.ptr(out), None, fl) == 0
        assert (out == c).all()


@pytest.mark.parametrize("logn", range(1, 16))
def test_lazy32_all_sizes(emu, orc, logn):
    n = 1 << logn
    for loge in (0, 2, 3, 6):
        _check(emu, orc, 0, Q17, n, loge)
    _check(emu, orc, 0, Q30, n, 0)


@pytest.mark.parametrize("logn", range(1, 16))
def test_small32_all_sizes(emu, orc, logn):
    # q < 2^22: csub-free forward butterflies, Montgomery pointwise with the factor folded into n^-1
    # and csub-free inverse butterflies, which need 2q*n <= 2^32 (that is what selects the policy)
    n = 1 << logn
    small = lambda q: (2 * q) << logn <= 2**32
    for loge in (0, 3, 4):
        if small(Q17):
            _check(emu, orc, 3, Q17, n, loge)
    if small(Q22):
        _check(emu, orc, 3, Q22, n, 0)
    else:  # the library must refuse the policy there (and pick Lazy32 on its own)
        a = orc.uniform(1, n, Q22)
        assert emu.emu_ntt(3, Q22, n, 0, 0, orc.ptr(a), None, orc.ptr(a.copy()), None, 0) == -2
        _check(emu, orc, -1, Q22, n, 0)
    _check(emu, orc, -1, Q17, n, 0)  # the policy the library picks for the reference's modulus


@pytest.mark.parametrize("logn", range(1, 15))
def test_64bit_policies_all_sizes(emu, orc, logn):
    n = 1 << logn
    _check(emu, orc, 1, Q62, n, 0)
    _check(emu, orc, 1, Q17, n, 5)
    _check(emu, orc, 2, Q63, n, 0)
    _check(emu, orc, 2, Q62, n, 3)


def test_plan_matches_oracle_tables(emu, orc):
    for q, n in [(Q17, 4), (Q17, 1024), (Q62, 512), (Q63, 64), (Q30, 2048)]:
        psi, ninv = np.zeros(1, np.uint64), np.zeros(1, np.uint64)
        r, ri = np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        assert emu.emu_plan(q, n, orc.ptr(psi), orc.ptr(ninv), orc.ptr(r), orc.ptr(ri)) == 0
        ro, rio, ninvo = orc.ntt_tables(q, n)
        assert (r == ro).all() and (ri == rio).all() and int(ninv[0]) == ninvo
        assert int(psi[0]) == orc.lib().orc_primitive_root_of_unity(q, 2 * n)
    # psi for the reference's modulus
    for n, want in [(4, 4096), (512, 19139), (1024, 61869), (4096, 6561), (16384, 9)]:
        psi = np.zeros(1, np.uint64)
        r = np.zeros(n, np.uint64)
        assert emu.emu_plan(Q17, n, orc.ptr(psi), orc.ptr(psi.copy()), orc.ptr(r), orc.ptr(r.copy())) == 0
        assert int(psi[0]) == want


def test_modmul_policies(emu):
    rng = np.random.default_rng(5)
    for kind, q in [(0, Q17), (0, Q30), (0, 7), (0, 17), (1, Q62), (1, Q17), (2, Q63), (2, Q62), (3, Q17), (3, Q22), (3, 17)]:
        cases = [(q - 1, q - 1), (0, q - 1), (1, q - 1), (q - 1, 1), (0, 0)]
        cases += [(int(x), int(y)) for x, y in zip(rng.integers(0, q, 5000, dtype=np.uint64), rng.integers(0, q, 5000, dtype=np.uint64))]
        for a, b in cases:
            assert emu.emu_modmul(kind, q, a, b) == a * b % q


@pytest.mark.parametrize("q,logn", [(Q17, 14), (Q17, 10), (Q22, 9), (417793, 12), (1032193, 11), (163841, 13)])
def test_small32_inverse_worst_case_growth(emu, orc, q, logn):
    # csub-free inverse: the sum path reaches 2q*n - 1 when every input is q-1 (and, for the polymul, 2q-1 after the
    # lazy pointwise product); must still equal the oracle at the largest (q, n) the policy accepts
    n = 1 << logn
    for fill in (q - 1, 0, 1):
        a = np.full(n, fill, dtype=np.uint64)
        out = np.empty(n, dtype=np.uint64)
        assert emu.emu_ntt(3, q, n, 0, 1, orc.ptr(a), None, orc.ptr(out), None, 0) == 0
        assert np.array_equal(out, orc.ntt(q, n, a, inverse=True))
        assert emu.emu_ntt(3, q, n, 0, 2, orc.ptr(a), orc.ptr(a), orc.ptr(out), None, 0) == 0
        assert np.array_equal(out, orc.rq_mul_batch(q, n, a, a))


@pytest.mark.parametrize("logn", [6, 7, 8, 9, 10, 13])  # pass 0 of the 32-per-thread shape holds >= 3 stages
def test_extprod_digit_transform_from_octet_table(emu, orc, logn):
    """xp_octet.cuh: the first three stages of a bit polynomial's transform read from the 256-entry table, then the
    ordinary passes -- against the oracle's NTT under both CRT primes of the torus path (torus.cuh)."""
    n = 1 << logn
    rng = np.random.default_rng(logn)
    cases = [rng.integers(0, 2, n, dtype=np.uint64), np.ones(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64)]
    one_hot = np.zeros(n, dtype=np.uint64)
    one_hot[n - 1] = 1
    cases.append(one_hot)
    for q in (0x7E90001, 0x7E00001):
        for bits in cases:
            out = np.zeros(n, dtype=np.uint64)
            assert emu.emu_xp_digit(q, n, orc.ptr(bits), orc.ptr(out)) == 0
            assert (out < 2**28).all()
            assert (out % np.uint64(q) == orc.ntt(q, n, bits)).all()


# ---- Fermat32: radix-4 butterflies for q = 65537 (modarith.cuh) ------------------------------------------------------
@pytest.mark.parametrize("logn", range(1, 16))
def test_fermat32_all_sizes(emu, orc, logn):
    # every degree, every coefficients-per-thread setting that changes the pass split (and with it which stages pair
    # up and which table slots hold parent * child products), against the oracle's radix-2 loop
    n = 1 << logn
    for loge in (0, 2, 3, 4, 5, 6):
        _check(emu, orc, 4, Q17, n, loge)
    # the policy is for the reference's modulus only
    a = orc.uniform(1, n, Q22 if (Q22 - 1) % (2 * n) == 0 else Q30)
    q_other = Q22 if (Q22 - 1) % (2 * n) == 0 else Q30
    assert emu.emu_ntt(4, q_other, n, 0, 0, orc.ptr(a), None, orc.ptr(a.copy()), None, 0) == -2


@pytest.mark.parametrize("logn", [2, 5, 9, 10, 11, 12, 13, 14, 15])
def test_fermat32_worst_case_ranges(emu, orc, logn):
    """The radix-4 forward lets the never-multiplied path grow by up to 1026q per layer and shifts differences by 8
    bits; the inverse shifts differences of sums that double every stage.  Extreme and structured inputs (all q-1,
    one-hot, alternating 0 / q-1, q-1 on one residue class of every power-of-two stride) drive those paths to their
    bounds; results must still equal the oracle."""
    n = 1 << logn
    q = Q17
    cases = [np.full(n, q - 1, dtype=np.uint64), np.zeros(n, dtype=np.uint64), np.ones(n, dtype=np.uint64)]
    alt = np.zeros(n, dtype=np.uint64)
    alt[::2] = q - 1
    cases.append(alt)
    cases.append((q - 1) - alt)
    stride = 1
    while stride < n:
        v = np.zeros(n, dtype=np.uint64)
        v[(np.arange(n) // stride) % 2 == 0] = q - 1
        cases.append(v)
        stride *= 4
    hot = np.zeros(n, dtype=np.uint64)
    hot[n - 1] = q - 1
    cases.append(hot)
    rng = np.random.default_rng(logn)
    cases.append(rng.integers(q - 4, q, n, dtype=np.uint64))
    out = np.empty(n, dtype=np.uint64)
    for loge in (0, 3):
        for a in cases:
            fa = orc.ntt(q, n, a)
            assert emu.emu_ntt(4, q, n, loge, 0, orc.ptr(a), None, orc.ptr(out), None, 0) == 0
            assert np.array_equal(out, fa)
            assert emu.emu_ntt(4, q, n, loge, 1, orc.ptr(a), None, orc.ptr(out), None, 0) == 0
            assert np.array_equal(out, orc.ntt(q, n, a, inverse=True))
            for b in (a, cases[0], cases[-1]):
                assert emu.emu_ntt(4, q, n, loge, 2, orc.ptr(a), orc.ptr(b), orc.ptr(out), None, 0) == 0
                assert np.array_equal(out, orc.rq_mul_batch(q, n, a, b))
                # evals given for either operand (canonical NTT values straight into the lazy pointwise product)
                assert emu.emu_ntt(4, q, n, loge, 2, orc.ptr(fa), orc.ptr(b), orc.ptr(out), None, 1) == 0
                assert np.array_equal(out, orc.rq_mul_batch(q, n, a, b))


def test_fermat32_is_what_the_library_picks(emu, orc):
    for n in (2, 4, 1024, 16384):
        _check(emu, orc, -1, Q17, n, 0)


def test_fermat32_unreduced_17_bit_inputs(emu, orc):
    """The shift-only first layer is sized for any 17-bit input word: coefficients in [q, 2^17) are outside the
    documented precondition (canonical input) but still give the residue's transform and product."""
    q = Q17
    rng = np.random.default_rng(17)
    for logn in (2, 6, 10, 12, 15):
        n = 1 << logn
        a = rng.integers(0, 1 << 17, n, dtype=np.uint64)
        a[0], a[-1] = (1 << 17) - 1, q
        b = np.full(n, (1 << 17) - 1, dtype=np.uint64)
        ar, br = a % np.uint64(q), b % np.uint64(q)
        out = np.empty(n, dtype=np.uint64)
        assert emu.emu_ntt(4, q, n, 0, 0, orc.ptr(a), None, orc.ptr(out), None, 0) == 0
        assert np.array_equal(out, orc.ntt(q, n, ar))
        assert emu.emu_ntt(4, q, n, 0, 2, orc.ptr(a), orc.ptr(b), orc.ptr(out), None, 0) == 0
        assert np.array_equal(out, orc.rq_mul_batch(q, n, ar, br))
