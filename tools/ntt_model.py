"""Python model of the register-blocked NTT pass structure used by csrc/ntt_core.cuh.

Validates the index / twiddle math against the reference loop (arith/src/ntt.rs:44-110 semantics) and
counts shared-memory bank conflicts of every exchange for a (LOGN, LOGE) configuration."""
import sys
from collections import Counter

Q = 65537


def split(LOGN, LOGE):
    P = -(-LOGN // LOGE)
    base, rem = divmod(LOGN, P)
    if P >= 3:  # pass 0 takes the remainder, every later pass LOGE stages (ntt_core.cuh: FRONT)
        return [LOGN - (P - 1) * LOGE] + [LOGE] * (P - 1)
    return [base + 1 if p < rem else base for p in range(P)]


def bitrev(i, b):
    return int(format(i, f"0{b}b")[::-1], 2) if b else 0


def tables(q, n):
    logn = n.bit_length() - 1
    k = 1
    while True:
        w = pow(k, (q - 1) // (2 * n), q)
        if pow(w, n, q) != 1:
            break
        k += 1
    roots = [pow(w, bitrev(i, logn), q) for i in range(n)]
    return roots, [pow(r, q - 2, q) for r in roots], pow(n, q - 2, q)


def ref_ntt(a, q, roots):
    n = len(a)
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


def ref_intt(a, q, roots_inv, n_inv):
    n = len(a)
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


def layout(LOGN, LOGE, gs, p, tid):
    """index of register slot e of thread tid in pass p's layout"""
    E, T = 1 << LOGE, 1 << (LOGN - LOGE)
    s0, g = sum(gs[:p]), gs[p]
    nL = LOGN - s0 - g
    out = []
    for qi in range(E >> g):
        u = tid + qi * T
        H, L = u >> nL, u & ((1 << nL) - 1)
        for r in range(1 << g):
            out.append(((H << (LOGN - s0)) | (r << nL) | L, H))
    return out


def model_fwd(a, q, roots, LOGN, LOGE, inverse=False, roots_inv=None, n_inv=None):
    gs = split(LOGN, LOGE)
    E, T = 1 << LOGE, 1 << (LOGN - LOGE)
    mem = list(a)
    order = range(len(gs)) if not inverse else reversed(range(len(gs)))
    for p in order:
        s0, g = sum(gs[:p]), gs[p]
        new = list(mem)
        for tid in range(T):
            lay = layout(LOGN, LOGE, gs, p, tid)
            x = [mem[i] for i, _ in lay]
            stages = range(g) if not inverse else reversed(range(g))
            for ls in stages:
                s = s0 + ls
                half = 1 << (g - 1 - ls)
                for qi in range(E >> g):
                    H = lay[qi << g][1]
                    for ru in range(1 << g):
                        if ru & half:
                            continue
                        rv = ru | half
                        twi = (1 << s) + (H << ls) + (ru >> (g - ls))
                        iu, iv = (qi << g) + ru, (qi << g) + rv
                        if not inverse:
                            U, V = x[iu], x[iv] * roots[twi] % q
                            x[iu], x[iv] = (U + V) % q, (U - V) % q
                        else:
                            U, V = x[iu], x[iv]
                            x[iu], x[iv] = (U + V) % q, (U - V) * roots_inv[twi] % q
            for (i, _), v in zip(lay, x):
                new[i] = v
        mem = new
    if inverse:
        mem = [v * n_inv % q for v in mem]
    return mem


def pad_rule(LOGN, LOGE, word_bytes=4):
    """(pad words per row, vectorised) as PadRule in csrc/ntt_kernels.cuh: 4-byte words with 32 coefficients per thread
    and a five-stage last pass pad FOUR words per 32 and access the last-pass layout (a thread's 32 consecutive words)
    with 128-bit instructions; everything else pads one word per 128 bytes."""
    gs = split(LOGN, LOGE)
    P = len(gs)
    nL = [LOGN - sum(gs[: p + 1]) for p in range(P)]
    vec = word_bytes == 4 and LOGE == 5 and P >= 2 and gs[-1] == LOGE and nL[P - 2] >= 5
    return (4 if vec else 1), vec


def conflicts(LOGN, LOGE, word_bytes=4):
    """worst bank-conflict degree of each pass layout (1 = conflict-free) under the kernels' padding (pad_rule; an
    8-byte access is served per half-warp, a 16-byte access per quarter-warp)"""
    gs = split(LOGN, LOGE)
    E, T = 1 << LOGE, 1 << (LOGN - LOGE)
    lanes, banks_n, pad_shift = (32, 32, 5) if word_bytes == 4 else (16, 16, 4)
    K, vec = pad_rule(LOGN, LOGE, word_bytes)
    pad = lambda i: i + K * (i >> pad_shift)
    worst = {}
    for p in range(len(gs)):
        w = 0
        row = vec and p == len(gs) - 1   # 128-bit accesses: 8 threads per phase, 4 consecutive words each
        step = 8 if row else lanes
        for t0 in range(0, T, step):
            lays = [layout(LOGN, LOGE, gs, p, t) for t in range(t0, min(t0 + step, T))]
            for e in range(0, E, 4 if row else 1):
                banks = Counter()
                for lay in lays:
                    if row:
                        base = pad(lay[e][0])
                        assert base % 4 == 0 and [pad(lay[e + j][0]) for j in range(4)] == [base + j for j in range(4)]
                        for j in range(4):
                            banks[(base + j) % banks_n] += 1
                    else:
                        banks[pad(lay[e][0]) % banks_n] += 1
                w = max(w, max(banks.values()))
        worst[p] = w
    return gs, worst


if __name__ == "__main__":
    import random
    for LOGN, LOGE in [(2, 2), (4, 4), (6, 3), (7, 4), (8, 4), (9, 5), (10, 5), (11, 4), (12, 4), (13, 5), (14, 5), (10, 4), (12, 5)]:
        n = 1 << LOGN
        roots, roots_inv, n_inv = tables(Q, n)
        a = [random.randrange(Q) for _ in range(n)]
        f = model_fwd(a, Q, roots, LOGN, LOGE)
        assert f == ref_ntt(a, Q, roots), (LOGN, LOGE)
        b = model_fwd(f, Q, roots, LOGN, LOGE, True, roots_inv, n_inv)
        assert b == a and b == ref_intt(f, Q, roots_inv, n_inv)
        print(LOGN, LOGE, "ok", conflicts(LOGN, LOGE))
