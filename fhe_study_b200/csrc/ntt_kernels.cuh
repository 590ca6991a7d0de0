// ntt_kernels.cuh -- batched negacyclic NTT / INTT / R_q polymul kernels (one polynomial per group of
// T = N/E threads, E coefficients per thread in registers, one shared-memory exchange between
// register-local passes).  Included by one translation unit per modular policy (ntt_inst_*.cu).
//
// Follows arith/src/ntt.rs:44-110 (NTT::ntt / NTT::intt) and arith/src/ring_nq.rs:564-607
// (mul / mul_mut: A = evals or ntt(a), B = evals or ntt(b), C = A.B pointwise, c = intt(C), the
// result keeps C as its cached evals).
#pragma once
#include <stdlib.h>
#include "common.cuh"
#include "ntt_core.cuh"

namespace fhe {

// Device-visible plan for one (q, n).  Passed BY VALUE (__grid_constant__): c_fwd / c_inv (the first
// 2^g0 <= 64 table entries, all that pass 0 needs) are then read straight from the constant bank.
template <class M> struct NttParams {
    M mod;
    const typename M::T *fwd;   // n entries: roots[i]     = psi^bitrev(i)      (arith/src/ntt.rs:133-147)
    const typename M::T *inv;   // n entries: roots_inv[i] = roots[i]^-1        (arith/src/ntt.rs:149-161)
    typename M::T ninv;         // n^-1                                        (arith/src/ntt.rs:27-30)
    typename M::T s_ninv;       // roots_inv[1] * n^-1
    typename M::T ninv_pw, s_ninv_pw;  // same, times the scale factor M::pw_mul leaves behind (polymul only)
    typename M::T c_fwd[64], c_inv[64];
};

enum NttMode { MODE_FWD = 0, MODE_INV = 1, MODE_MUL = 2 };
enum MulFlags { A_IS_EVALS = 1, B_IS_EVALS = 2, B_BROADCAST = 4 };  // B_BROADCAST: b is ONE polynomial, used for every product

template <int LOGN, int LOGE> struct KernelGeom {
    typedef NttShape<LOGN, LOGE> S;
#ifndef FHE_NTT_MIN_CT
#define FHE_NTT_MIN_CT 128
#endif
    static constexpr int CT = S::T > FHE_NTT_MIN_CT ? S::T : FHE_NTT_MIN_CT;   // threads per CTA
    static constexpr int PPC = CT / S::T;                // polynomials per CTA
    template <int WB> __host__ __device__ static constexpr int padn() { return S::N + (S::N >> (WB == 4 ? 5 : 4)); }  // padded words per polynomial in smem
    static constexpr int PADN = padn<4>();
};
__host__ __device__ constexpr int pad_idx(int i) { return i + (i >> 5); }
// one pad word per 128 bytes: every 32 words of 4 bytes, every 16 words of 8 bytes
template <int WB> __host__ __device__ constexpr int pad_idx_w(int i) { return i + (i >> (WB == 4 ? 5 : 4)); }

template <int T> __device__ __forceinline__ void group_sync() {
    if (T <= 32) __syncwarp(); else __syncthreads();
}

// An exchange between ADJACENT passes a < b touches, per thread, a closed set of partners: the threads that share
// the s0(a) high bits H and the nL(b) low bits of the butterfly-group index.  When g(a) == g(b) the g bits that
// vary sit at [nL(b), nL(a)) in both layouts; with nL(a) <= 5 they are lane bits, so every partner is in the same
// warp and __syncwarp orders the exchange (always the case for the LAST two passes once they are equal).
template <int LOGN, int LOGE, int FROM, int TO> struct ExchScope {
    typedef NttShape<LOGN, LOGE> S;
    static constexpr int A = FROM < TO ? FROM : TO, B = FROM < TO ? TO : FROM;
    static constexpr bool in_warp = S::T <= 32 || (B == A + 1 && S::g(A) == S::g(B) && S::nL(A) <= 5 && S::T % 32 == 0);
};

// registers (layout of pass FROM) -> shared -> registers (layout of pass TO).
// LEAD = false drops the barrier in front of the writes: allowed when this thread's previous access to `sm` was
// the read side of an exchange INTO layout FROM (it then overwrites exactly the words it read itself, so there is
// no other reader to wait for) -- i.e. for every exchange of a chain but the first.
template <class M, int LOGN, int LOGE, int FROM, int TO, bool LEAD = true>
__device__ __forceinline__ void exchange(typename M::W (&x)[1 << LOGE], typename M::W *sm, int tid) {
    typedef NttShape<LOGN, LOGE> S;
#ifdef FHE_NTT_FULL_SYNC  // the conservative scheme (two CTA-wide barriers per exchange), kept for A/B measurements
    group_sync<S::T>();
#else
    if constexpr (LEAD) group_sync<S::T>();  // earlier readers of sm are done
#endif
    // pos(p, tid, e) = pos(p, tid, 0) | pos(p, 0, e) on disjoint bits (tid < T = 2^k and e >> g select different bits
    // of the group index), and i + (i >> 5) is additive over disjoint bits: thread base + compile-time offset, so
    // every STS/LDS below takes an immediate offset instead of per-element LOP3/LEA address arithmetic.
    constexpr int WB = (int)sizeof(typename M::W);
    const int bf = pad_idx_w<WB>(S::pos(FROM, tid, 0)), bt = pad_idx_w<WB>(S::pos(TO, tid, 0));
#pragma unroll
    for (int e = 0; e < S::E; e++) sm[bf + pad_idx_w<WB>(S::pos(FROM, 0, e))] = x[e];
#ifdef FHE_NTT_FULL_SYNC
    group_sync<S::T>();
#else
    if constexpr (ExchScope<LOGN, LOGE, FROM, TO>::in_warp) __syncwarp(); else __syncthreads();
#endif
#pragma unroll
    for (int e = 0; e < S::E; e++) x[e] = sm[bt + pad_idx_w<WB>(S::pos(TO, 0, e))];
}

template <class M, int LOGN, int LOGE, int PASS = 0, bool FIRST = true>
__device__ __forceinline__ void fwd_chain(typename M::W (&x)[1 << LOGE], typename M::W *sm, int tid, const M &m,
                                          const TwSrc<M> &tw) {
    typedef NttShape<LOGN, LOGE> S;
    if constexpr (PASS > 0) exchange<M, LOGN, LOGE, PASS - 1, PASS, FIRST>(x, sm, tid);
    fwd_pass<M, LOGN, LOGE, PASS>(x, tid, m, tw);
    if constexpr (PASS + 1 < S::P) fwd_chain<M, LOGN, LOGE, PASS + 1, (FIRST && PASS == 0)>(x, sm, tid, m, tw);
}
template <class M, int LOGN, int LOGE, int PASS, bool FIRST = true>
__device__ __forceinline__ void inv_chain(typename M::W (&x)[1 << LOGE], typename M::W *sm, int tid, const M &m,
                                          const TwSrc<M> &tw, typename M::T ninv, typename M::T s_ninv) {
    inv_pass<M, LOGN, LOGE, PASS>(x, tid, m, tw, ninv, s_ninv);
    if constexpr (PASS > 0) {
        exchange<M, LOGN, LOGE, PASS, PASS - 1, FIRST>(x, sm, tid);
        inv_chain<M, LOGN, LOGE, PASS - 1, false>(x, sm, tid, m, tw, ninv, s_ninv);
    }
}

// Polynomial data is touched once.  STREAM reads it without allocating in L1 and stores it evict-first, so that L1
// keeps the twiddle tables (at N >= 4096 the per-thread twiddles of the last pass are as many bytes as the
// polynomial).  Measured (q = 65537): polymul N=4096/8192/16384 +1/+2/+3 %, but the transform-only kernels LOSE
// 2..7 % -- hence on for the polymul at N >= 4096 only (ntt_stream()).
#ifndef FHE_NTT_STREAM
#define FHE_NTT_STREAM 1
#endif
__host__ __device__ constexpr bool ntt_stream(int logn, int mode) { return FHE_NTT_STREAM && mode == 2 /* MODE_MUL */ && logn >= 12; }
#ifndef FHE_NTT_PREFETCH
#define FHE_NTT_PREFETCH 0  // measured slower (N=8192 polymul 0.47 -> 0.38 of HBM peak): kept for reference only
#endif
__host__ __device__ constexpr bool ntt_prefetch(int logn, int mode) { return FHE_NTT_PREFETCH && mode == 2 && logn >= 11; }
// one prefetch per 128-byte line of an N-coefficient polynomial, spread over its T threads
template <int N, int T> __device__ __forceinline__ void prefetch_poly_l2(const u64 *g, int tid) {
    constexpr int LINES = N / 16;
#pragma unroll
    for (int i = 0; i < (LINES + T - 1) / T; i++) {
        const int line = tid + i * T;
        if (LINES % T == 0 || line < LINES) asm volatile("prefetch.global.L2 [%0];" ::"l"(g + (size_t)line * 16));
    }
}
template <bool STREAM> __device__ __forceinline__ u64 ld_poly(const u64 *p) {
    if constexpr (STREAM) {
        u64 v;
        asm("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
        return v;
    } else {
        return __ldg(p);
    }
}
template <bool STREAM> __device__ __forceinline__ void st_poly(u64 *p, u64 v) {
    if constexpr (STREAM) __stcs(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
    else *p = v;
}

// global (coalesced, layout of pass 0) -> registers in the layout of pass TO
template <class M, int LOGN, int LOGE, int TO, bool STREAM = false>
__device__ __forceinline__ void load_poly(typename M::W (&x)[1 << LOGE], const u64 *__restrict__ g, bool valid,
                                          typename M::W *sm, int tid) {
    typedef NttShape<LOGN, LOGE> S;
    const u64 *gt = g + S::pos(0, tid, 0);
#pragma unroll
    for (int e = 0; e < S::E; e++) x[e] = valid ? M::load(ld_poly<STREAM>(gt + S::pos(0, 0, e))) : (typename M::W)0;
    if constexpr (TO != 0) exchange<M, LOGN, LOGE, 0, TO>(x, sm, tid);
}
// registers in the layout of pass FROM -> global (coalesced).  LEAD as in exchange().
template <class M, int LOGN, int LOGE, int FROM, bool LEAD = true, bool STREAM = false>
__device__ __forceinline__ void store_poly(typename M::W (&x)[1 << LOGE], u64 *__restrict__ g, bool valid,
                                           typename M::W *sm, int tid) {
    typedef NttShape<LOGN, LOGE> S;
    if constexpr (FROM != 0) exchange<M, LOGN, LOGE, FROM, 0, LEAD>(x, sm, tid);
    if (valid) {
        u64 *gt = g + S::pos(0, tid, 0);
#pragma unroll
        for (int e = 0; e < S::E; e++) st_poly<STREAM>(gt + S::pos(0, 0, e), M::store(x[e]));
    }
}

// Polymul with NTT(a) parked in shared memory while b is transformed (instead of 32 more live registers): lets more
// CTAs be resident.  Only for one-warp transforms on 32-bit words.  FHE_A_SMEM_MINB = resident CTAs asked of ptxas.
#ifndef FHE_A_SMEM_MINB
#define FHE_A_SMEM_MINB 0
#endif
template <class M, int LOGN, int LOGE, int MODE> struct ASmem {
    static constexpr bool on = FHE_A_SMEM_MINB > 0 && MODE == MODE_MUL && sizeof(typename M::W) == 4 &&
                               NttShape<LOGN, LOGE>::T <= 32 && LOGN == 10;
#ifndef FHE_MUL_MINB
#define FHE_MUL_MINB 5
#endif
    // 0 = no minimum (ptxas' own choice; an explicit 1 makes it spend up to 65536/CT registers and lose occupancy).
    // The 32-bit polymul on 128-thread CTAs asks for five resident CTAs (<= 102 registers, no spills): measured
    // 0.775 of HBM peak at N=1024 against 0.74 with ptxas' unconstrained 139 registers (three CTAs).
    // measured (q = 65537, hbm fraction before -> after): polymul N=1024 0.74 -> 0.78 (CT=128, five CTAs); NTT N=16384
    // 0.47 -> 0.57 (CT=512, two CTAs of 64 registers instead of one of ~100).  CT=256 (N=8192): with immediate-offset
    // exchanges ptxas settles for 48 registers on the forward transform and serialises its loads (0.71); a bound of
    // three CTAs lets it spend 80 and hoist them (0.78) -- the opposite of what the bound did to the older code
    static constexpr int CT_ = KernelGeom<LOGN, LOGE>::CT;
    static constexpr bool W32 = sizeof(typename M::W) == 4;
#ifndef FHE_NTT_MINB_256
#define FHE_NTT_MINB_256 3
#endif
#ifndef FHE_NTT_MINB_512
#define FHE_NTT_MINB_512 2
#endif
#ifndef FHE_INV_MINB_256
#define FHE_INV_MINB_256 3
#endif
#ifndef FHE_MUL_MINB_256
#define FHE_MUL_MINB_256 3
#endif
#ifndef FHE_MUL_MINB_E16   // 128-thread polymul CTAs with 16 coefficients per thread (N=2048)
#define FHE_MUL_MINB_E16 8  // 63 registers, no spills: N=2048 polymul 0.652 (5 CTAs) -> 0.677 (6) -> 0.690 (7) -> 0.697 (8) -> 0.701 (9, spills)
#endif
#ifndef FHE_MUL64_MINB
#define FHE_MUL64_MINB 0
#endif
#ifndef FHE_NTT64_MINB
#define FHE_NTT64_MINB 0
#endif
    static constexpr int minb = on ? FHE_A_SMEM_MINB
                                : !W32 ? (CT_ == 128 ? (MODE == MODE_MUL ? FHE_MUL64_MINB : FHE_NTT64_MINB) : 0)
                                : MODE == MODE_MUL ? (CT_ == 128 ? (LOGE == 4 ? FHE_MUL_MINB_E16 : FHE_MUL_MINB) : CT_ == 64 ? 2 * FHE_MUL_MINB : CT_ == 256 ? FHE_MUL_MINB_256 : 0)
                                : (CT_ == 256 ? (MODE == MODE_INV ? FHE_INV_MINB_256 : FHE_NTT_MINB_256)
                                   : CT_ == 512 ? FHE_NTT_MINB_512 : 0);
};

template <class M, int LOGN, int LOGE, int MODE>
__global__ void __launch_bounds__(KernelGeom<LOGN, LOGE>::CT, ASmem<M, LOGN, LOGE, MODE>::minb)
ntt_kernel(const __grid_constant__ NttParams<M> P, const u64 *__restrict__ a, const u64 *__restrict__ b,
           u64 *__restrict__ c, u64 *__restrict__ c_evals, size_t batch, int flags, int pf_dist) {
    typedef NttShape<LOGN, LOGE> S;
    typedef KernelGeom<LOGN, LOGE> G;
    typedef typename M::W W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int slot = threadIdx.x / S::T, tid = threadIdx.x % S::T;
    W *sm = reinterpret_cast<W *>(smem_raw) + (size_t)slot * G::template padn<sizeof(W)>();
    const size_t poly = (size_t)blockIdx.x * G::PPC + slot;
    const bool valid = poly < batch;
    // slots past the end of a ragged batch read the last polynomial (and store nothing): unconditional loads, no
    // zero-filled registers
    const size_t off_ld = (valid ? poly : batch - 1) * S::N;
    const size_t off = poly * S::N;
    const M &m = P.mod;
    constexpr int LAST = S::P - 1;
    W x[S::E];

    if constexpr (MODE == MODE_FWD) {
        const TwSrc<M> tw = {P.c_fwd, P.fwd};
        load_poly<M, LOGN, LOGE, 0>(x, a + off_ld, true, sm, tid);
        fwd_chain<M, LOGN, LOGE>(x, sm, tid, m, tw);
#pragma unroll
        for (int e = 0; e < S::E; e++) x[e] = m.fwd_canon(x[e]);
        store_poly<M, LOGN, LOGE, LAST, false>(x, c + off, valid, sm, tid);  // last smem access: own reads in layout LAST
    } else if constexpr (MODE == MODE_INV) {
        const TwSrc<M> tw = {P.c_inv, P.inv};
        load_poly<M, LOGN, LOGE, LAST>(x, a + off_ld, true, sm, tid);
        inv_chain<M, LOGN, LOGE, LAST>(x, sm, tid, m, tw, P.ninv, P.s_ninv);
#pragma unroll
        for (int e = 0; e < S::E; e++) x[e] = m.canon2(x[e]);
        store_poly<M, LOGN, LOGE, 0>(x, c + off, valid, sm, tid);
    } else {
        const TwSrc<M> twf = {P.c_fwd, P.fwd};
        const TwSrc<M> twi = {P.c_inv, P.inv};
        constexpr bool PARK = ASmem<M, LOGN, LOGE, MODE>::on;
        constexpr bool ST = ntt_stream(LOGN, MODE);
        if constexpr (ntt_prefetch(LOGN, MODE)) {
            // Pull b into L2 while a is loaded and transformed, and both operands of the polynomial this SM slot
            // will most likely process next (pf_dist = CTAs resident on the whole GPU, CTAs being dispatched in
            // order): the few resident warps of the large degrees cannot hide an HBM round trip per operand.
            const bool b_own = !(flags & B_BROADCAST);
            if (valid && b_own) prefetch_poly_l2<S::N, S::T>(b + off, tid);
            const size_t nxt = poly + (size_t)pf_dist * G::PPC;
            if (pf_dist > 0 && nxt < batch) {
                prefetch_poly_l2<S::N, S::T>(a + nxt * S::N, tid);
                if (b_own) prefetch_poly_l2<S::N, S::T>(b + nxt * S::N, tid);
            }
        }
        W A[PARK ? 1 : S::E];
        W *sA = reinterpret_cast<W *>(smem_raw) + (size_t)G::PPC * G::template padn<sizeof(W)>() + (size_t)slot * S::N;  // PARK only
        // both operands run through ONE copy of the forward-transform code (the fully unrolled transform is
        // the bulk of the kernel's instruction footprint; see profiles/: no_instruction stalls)
#pragma unroll 1
        for (int op = 0; op < 2; op++) {
            const u64 *src = op == 0 ? a + off_ld : (flags & B_BROADCAST) ? b : b + off_ld;
            if (flags & (op == 0 ? A_IS_EVALS : B_IS_EVALS)) {
                load_poly<M, LOGN, LOGE, LAST, ST>(x, src, true, sm, tid);
            } else {
                load_poly<M, LOGN, LOGE, 0, ST>(x, src, true, sm, tid);
                fwd_chain<M, LOGN, LOGE>(x, sm, tid, m, twf);
#pragma unroll
                for (int e = 0; e < S::E; e++) x[e] = m.fwd_out(x[e]);
            }
            if (op == 0) {
                if constexpr (PARK) {
#pragma unroll
                    for (int e = 0; e < S::E; e++) sA[e * S::T + tid] = x[e];
                } else {
#pragma unroll
                    for (int e = 0; e < S::E; e++) A[e] = x[e];
                }
            }
        }
        if constexpr (PARK) {
#pragma unroll
            for (int e = 0; e < S::E; e++) x[e] = m.pw_mul(sA[e * S::T + tid], x[e]);
        } else {
#pragma unroll
            for (int e = 0; e < S::E; e++) x[e] = m.pw_mul(A[e], x[e]);
        }
        if (c_evals != nullptr) {  // ring_nq.rs:606 -- the product keeps its evals
            W ev[S::E];
#pragma unroll
            for (int e = 0; e < S::E; e++) ev[e] = m.pw_evals(x[e]);
            // b's last smem access (end of its chain, or load_poly<LAST>) was this thread's reads in layout LAST
            store_poly<M, LOGN, LOGE, LAST, false, ST>(ev, c_evals + off, valid, sm, tid);
            if constexpr (S::P > 1) group_sync<S::T>();  // the evals store read layout 0: foreign words
        }
        inv_chain<M, LOGN, LOGE, LAST, false>(x, sm, tid, m, twi, P.ninv_pw, P.s_ninv_pw);
#pragma unroll
        for (int e = 0; e < S::E; e++) x[e] = m.canon2(x[e]);
        store_poly<M, LOGN, LOGE, 0, true, ST>(x, c + off, valid, sm, tid);
    }
}

template <class M, int LOGN, int LOGE, int MODE>
int launch_one(const NttParams<M> &P, const u64 *a, const u64 *b, u64 *c, u64 *c_evals, size_t batch, int flags,
               cudaStream_t st) {
    typedef KernelGeom<LOGN, LOGE> G;
    auto kern = ntt_kernel<M, LOGN, LOGE, MODE>;
    const size_t smem = (size_t)G::PPC * (G::template padn<sizeof(typename M::W)>() + (ASmem<M, LOGN, LOGE, MODE>::on ? G::S::N : 0)) * sizeof(typename M::W);
    if (smem > 48 * 1024) {  // opt in once per device (and per instantiation)
        static unsigned long long done_mask = 0;
        int dev = 0;
        FHE_CUDA_OK(cudaGetDevice(&dev));
        if (!((done_mask >> (dev & 63)) & 1ull)) {
            FHE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            done_mask |= 1ull << (dev & 63);
        }
    }
    {   // tuning knob: FHE_NTT_CARVE = preferred shared-memory carve-out in percent (caps the resident CTAs, leaves
        // the rest of the 256 KB to L1 for the twiddle tables); unset = the driver's choice
        static int carve_done = 0;
        if (!carve_done) {
            if (const char *e = getenv("FHE_NTT_CARVE"))
                FHE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
            carve_done = 1;
        }
    }
    const size_t grid = (batch + G::PPC - 1) / G::PPC;
    FHE_REQUIRE(grid <= 0x7fffffffull, "batch too large for one launch");
    int pf_dist = 0;
    if constexpr (ntt_prefetch(LOGN, MODE)) {  // CTAs resident on the whole device (cached per device)
        static int resident[64] = {0};
        int dev = 0;
        FHE_CUDA_OK(cudaGetDevice(&dev));
        if (resident[dev & 63] == 0) {
            int per_sm = 0;
            FHE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G::CT, smem));
            resident[dev & 63] = (per_sm > 0 ? per_sm : 1) * num_sms();
        }
        pf_dist = resident[dev & 63];
    }
    kern<<<(unsigned)grid, G::CT, smem, st>>>(P, a, b, c, c_evals, batch, flags, pf_dist);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

template <class M, int LOGN, int LE>
int launch_modes(int mode, const NttParams<M> &P, const u64 *a, const u64 *b, u64 *c, u64 *c_evals, size_t batch,
                 int flags, cudaStream_t st) {
    switch (mode) {
        case MODE_FWD: return launch_one<M, LOGN, LE, MODE_FWD>(P, a, b, c, c_evals, batch, flags, st);
        case MODE_INV: return launch_one<M, LOGN, LE, MODE_INV>(P, a, b, c, c_evals, batch, flags, st);
        default: return launch_one<M, LOGN, LE, MODE_MUL>(P, a, b, c, c_evals, batch, flags, st);
    }
}

// `loge` must be a value ntt_loge_supported() accepts for (M, logn): the default LogE<M>::of(logn), or, for
// the tunable degrees, one of the alternatives instantiated below.
// tunable degrees of the 32-bit policies: which coefficients-per-thread settings are instantiated
constexpr bool ntt_alt_loge(int logn, int loge, int wbytes) {
    return wbytes == 4 && ((logn >= 10 && logn <= 12 && loge >= 3 && loge <= 5) || (logn >= 13 && logn <= 14 && loge >= 4 && loge <= 5));
}
template <class M, int LOGN>
int launch_logn(int loge, int mode, const NttParams<M> &P, const u64 *a, const u64 *b, u64 *c, u64 *c_evals,
                size_t batch, int flags, cudaStream_t st) {
    constexpr int DEF = LogE<M>::of(LOGN);
    constexpr int WB = (int)sizeof(typename M::W);
    if constexpr (ntt_alt_loge(LOGN, 3, WB) && DEF != 3) {
        if (loge == 3) return launch_modes<M, LOGN, 3>(mode, P, a, b, c, c_evals, batch, flags, st);
    }
    if constexpr (ntt_alt_loge(LOGN, 4, WB) && DEF != 4) {
        if (loge == 4) return launch_modes<M, LOGN, 4>(mode, P, a, b, c, c_evals, batch, flags, st);
    }
    if constexpr (ntt_alt_loge(LOGN, 5, WB) && DEF != 5) {
        if (loge == 5) return launch_modes<M, LOGN, 5>(mode, P, a, b, c, c_evals, batch, flags, st);
    }
    if (loge != DEF) {
        set_error("internal: unsupported coefficients-per-thread setting");
        return -1;
    }
    return launch_modes<M, LOGN, DEF>(mode, P, a, b, c, c_evals, batch, flags, st);
}
template <class M> bool ntt_loge_supported(int logn, int loge) {
    if (loge == LogE<M>::of(logn)) return true;
    return ntt_alt_loge(logn, loge, (int)sizeof(typename M::W));
}

template <class M>
int launch_ntt(int logn, int loge, int mode, const NttParams<M> &P, const u64 *a, const u64 *b, u64 *c, u64 *c_evals,
               size_t batch, int flags, cudaStream_t st) {
    constexpr int MAXLOGN = sizeof(typename M::W) == 4 ? 15 : 14;  // one CTA's shared memory holds the polynomial
    if (logn < 1 || logn > MAXLOGN) {
        set_error("unsupported ring degree for this modulus width (n must be 2..2^15 for q<2^30, 2..2^14 otherwise)");
        return -1;
    }
    switch (logn) {
#define FHE_CASE(L) case L: return launch_logn<M, L>(loge, mode, P, a, b, c, c_evals, batch, flags, st);
        FHE_CASE(1) FHE_CASE(2) FHE_CASE(3) FHE_CASE(4) FHE_CASE(5) FHE_CASE(6) FHE_CASE(7) FHE_CASE(8)
        FHE_CASE(9) FHE_CASE(10) FHE_CASE(11) FHE_CASE(12) FHE_CASE(13) FHE_CASE(14)
#undef FHE_CASE
        case 15:
            if constexpr (sizeof(typename M::W) == 4)
                return launch_logn<M, 15>(loge, mode, P, a, b, c, c_evals, batch, flags, st);
    }
    set_error("unsupported ring degree");
    return -1;
}

}  // namespace fhe
