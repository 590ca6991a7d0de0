"""A SECOND, independent restatement of the reference's hot path in pure Python big integers -- test infrastructure
only.  It follows the Rust sources directly (file:line per function) and shares no code with oracle/fhe_oracle.c:
no 64-bit wrapping tricks (everything is computed over Python ints and reduced where Rust truncates), plain nested
loops, Python floats (IEEE doubles, like f64) for the f64 steps.  tests/test_pyref_crosscheck.py compares the C oracle
with it on small cases; it is the only extra pin available without a Rust toolchain."""
import math

M64 = 1 << 64


def wrap_i64(v):  # `as i64` of a wider integer (ring_n.rs:317: `*x as i64`)
    v &= M64 - 1
    return v - M64 if v >= 1 << 63 else v


def rust_round(x):  # f64::round: half away from zero
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def f64_as_i64(x):  # saturating cast, NaN -> 0
    if x != x:
        return 0
    return max(-(1 << 63), min((1 << 63) - 1, int(x)))


# ---- arith/src/zq.rs ---------------------------------------------------------------------------------------------
def zq_from_f64_exact(q, e):  # zq.rs:32-40: ((r % q) + q) % q with Rust's truncated (sign of the dividend) remainder
    r = f64_as_i64(rust_round(e))
    if r < 0 or r >= q:
        rem = abs(r) % q
        rem = -rem if r < 0 else rem
        r = (rem + q) % q          # rem + q > 0 because |rem| < q
    return r


# ---- arith/src/ntt.rs ----------------------------------------------------------------------------------------------
def ntt_tables(q, n):  # ntt.rs:115-161
    k = 1
    while True:
        w = pow(k, (q - 1) // (2 * n), q)
        if pow(w, n, q) != 1:
            break
        k += 1
    bits = n.bit_length() - 1
    rev = lambda i: int(format(i, f"0{bits}b")[::-1], 2) if bits else 0
    roots = [pow(w, rev(i), q) for i in range(n)]
    return roots, [pow(r, q - 2, q) for r in roots], pow(n, q - 2, q)


def ntt(q, n, a):  # ntt.rs:44-73
    roots, _, _ = ntt_tables(q, n)
    r = list(a)
    t, m = n // 2, 1
    while m < n:
        k = 0
        for i in range(m):
            S = roots[m + i]
            for j in range(k, k + t):
                U, V = r[j], r[j + t] * S % q
                r[j], r[j + t] = (U + V) % q, (U - V) % q
            k += 2 * t
        t //= 2
        m *= 2
    return r


def intt(q, n, a):  # ntt.rs:78-110
    _, roots_inv, n_inv = ntt_tables(q, n)
    r = list(a)
    t, m = 1, n // 2
    while m > 0:
        k = 0
        for i in range(m):
            S = roots_inv[m + i]
            for j in range(k, k + t):
                U, V = r[j], r[j + t]
                r[j], r[j + t] = (U + V) % q, (U - V) * S % q
            k += 2 * t
        t *= 2
        m //= 2
    return [x * n_inv % q for x in r]


def rq_mul(q, n, a, b):  # ring_nq.rs:586-607
    A, B = ntt(q, n, a), ntt(q, n, b)
    return intt(q, n, [x * y % q for x, y in zip(A, B)])


def rq_mul_schoolbook(q, n, a, b):
    """what an Rq product IS (negacyclic convolution mod q): cross-checks the NTT path itself"""
    c = [0] * n
    for i in range(n):
        for j in range(n):
            if i + j < n:
                c[i + j] += a[i] * b[j]
            else:
                c[i + j - n] -= a[i] * b[j]
    return [x % q for x in c]


# ---- arith/src/ring_torus.rs, torus.rs --------------------------------------------------------------------------------
def tn_mul(n, a, b):  # ring_torus.rs:266-298
    res = [0] * (2 * n - 1)
    for i in range(n):
        for j in range(n):
            res[i + j] = (res[i + j] + a[i] * b[j]) % (1 << 128)  # u128 accumulator
    for i in range(n, 2 * n - 1):
        res[i - n] = (res[i - n] - res[i]) % (1 << 128)        # wrapping_sub
    return [x % M64 for x in res[:n]]                              # `as u64`


def t64_decompose(x, l=64):  # torus.rs:43-52 (beta = 2): MSB first
    return [(x >> i) & 1 for i in range(l - 1, -1, -1)]


def tn_decompose(n, p, l=64):  # ring_torus.rs:67-77: poly j = digit j of every coefficient
    digs = [t64_decompose(c, l) for c in p]
    return [[digs[c][j] for c in range(n)] for j in range(l)]


def left_rotate(n, p, h):  # ring_torus.rs:118-132
    h %= n
    return p[h:] + [(-c) % M64 for c in p[:h]]


# ---- tfhe/src/tggsw.rs, tglwe.rs --------------------------------------------------------------------------------------
def tglwe_mul_tn(n, k, ct, p):  # tglwe.rs:182-194: every component times p
    return [tn_mul(n, ct[c], p) for c in range(k + 1)]


def tglwe_add(a, b):
    return [[(x + y) % M64 for x, y in zip(pa, pb)] for pa, pb in zip(a, b)]


def tglwe_sub(a, b):
    return [[(x - y) % M64 for x, y in zip(pa, pb)] for pa, pb in zip(a, b)]


def external_product(n, k, tggsw, ct):
    """tggsw.rs:45-62.  tggsw[i][j] = TGLWE (k+1 polys) of TGLev i, level j; ct = k+1 polys (mask then body)"""
    acc = [[0] * n for _ in range(k + 1)]
    for i in range(k + 1):
        digits = tn_decompose(n, ct[i])                     # decompose(2, 64)
        for j in range(64):
            acc = tglwe_add(acc, tglwe_mul_tn(n, k, tggsw[i][j], digits[j]))   # TGLev * Vec<Tn> (tggsw.rs:139-149)
    return acc


def cmux(n, k, tggsw, ct1, ct2):  # tggsw.rs:39-41
    return tglwe_add(ct1, external_product(n, k, tggsw, tglwe_sub(ct2, ct1)))


# ---- tfhe/src/tlwe.rs, tlev.rs -----------------------------------------------------------------------------------------
def key_switch(kn_in, kn_out, l, ksk, ct):
    """tlwe.rs:101-112; ksk[i][j] = TLWE (kn_out + 1 words); ct = kn_in + 1 words"""
    rhs = [0] * (kn_out + 1)
    for i in range(kn_in):
        digits = t64_decompose(ct[i], l)
        for j in range(l):                                 # TLev * Vec<T64> (tlev.rs:95-105)
            for x in range(kn_out + 1):
                rhs[x] = (rhs[x] + ksk[i][j][x] * digits[j]) % M64
    lhs = [0] * kn_out + [ct[kn_in]]
    return [(a - b) % M64 for a, b in zip(lhs, rhs)]


def mod_switch(x, q2):  # torus.rs:58-66
    assert q2 & (q2 - 1) == 0
    return x >> (64 - (q2.bit_length() - 1))


def sample_extraction(n, k, ct, h):  # tglwe.rs:89-115
    a = []
    for i in range(k):
        for j in range(n):
            a.append(ct[i][h - j] if j <= h else (-ct[i][n + h - j]) % M64)
    return a + [ct[k][h]]


def bootstrapping_as_executed(n, k, ksk, table, c, c_kn):  # tlwe.rs:121-161 (the CMux loop never runs)
    b = mod_switch(c[c_kn], k * n)
    rotated = [left_rotate(n, p, b) for p in table]
    ext = sample_extraction(n, k, rotated, 0)
    return key_switch(k * n, k * n, 64, ksk, ext)


# ---- arith/src/ring_n.rs, bfv/src/lib.rs --------------------------------------------------------------------------------
def naive_mul(n, a, b):  # ring_n.rs:307-320: linear convolution in i128, then `as i64`
    res = [0] * (2 * n - 1)
    for i in range(n):
        for j in range(n):
            res[i + j] += a[i] * b[j]
    return [wrap_i64(x) for x in res]


def fold(q, n, p):  # ring_nq.rs:132-141 with Zq::sub
    p = list(p)
    for i in range(n, len(p)):
        p[i - n] = (p[i - n] - p[i]) % q
    return p[:n]


def mul_div_round(q, n, v, num, den):  # ring_n.rs:130-138 -> Rq::from_vec_f64 (ring_nq.rs:160-163)
    return fold(q, n, [zq_from_f64_exact(q, (float(num) * float(e)) / float(den)) for e in v])


def bfv_tensor(q, n, t, a, b):  # lib.rs:59-85; a, b = (poly, poly)
    c0 = naive_mul(n, a[0], b[0])
    c1 = [wrap_i64(x + y) for x, y in zip(naive_mul(n, a[0], b[1]), naive_mul(n, a[1], b[0]))]
    c2 = naive_mul(n, a[1], b[1])
    return [mul_div_round(q, n, c, t, q) for c in (c0, c1, c2)]


def bfv_relinearize_204(q, n, pq, rlk, c0, c1, c2):  # lib.rs:251-271
    p = pq // q
    r0 = mul_div_round(q, n, naive_mul(n, c2, rlk[0]), 1, p)
    r1 = mul_div_round(q, n, naive_mul(n, c2, rlk[1]), 1, p)
    return [(x + y) % q for x, y in zip(c0, r0)], [(x + y) % q for x, y in zip(c1, r1)]


def bfv_mul(q, n, t, pq, rlk, a, b):  # lib.rs:87-90
    return bfv_relinearize_204(q, n, pq, rlk, *bfv_tensor(q, n, t, a, b))
