// tn_fused.cu -- fused Tn * Tn (arith/src/ring_torus.rs:266-298: the exact negacyclic product in
// Z_{2^64}[X]/(X^N+1)): one HBM round trip per product (read a and b, write c).  Same exact arithmetic as the
// unfused path in torus_kernels.cu -- both operands in four 16-bit limbs, two NTT primes below 2^27, the four
// weight classes 2^(16w) (only those reach the low 64 bits) bounded by 4*N*2^32 < P/2, centred CRT lift,
// recombination mod 2^64 -- but limb split, 16 forward transforms, the limb convolution, 8 inverse transforms
// and the CRT all stay on chip.
//
// A CTA of 256 threads runs SLOTS = 256/T concurrent transforms (T = N/32 threads each) = G = SLOTS/8 products
// per round: slot s = product s/8, operand (s%8)/4, limb s%4.  Per prime: forward round (csub-free butterflies,
// values brought below 2^28 by one shift + IMAD), pointwise limb convolution into 64-bit sums (< 4 * 2^56) reduced
// by one Barrett step, inverse transforms on the four class slots of every product.
#include "../../include/fhe_b200.h"
#include "ntt_kernels.cuh"
#include "runtime.cuh"
#include "torus.cuh"

namespace fhe {

template <int LOGN> struct TnGeom {
    static constexpr int N = 1 << LOGN;
    static constexpr int LOGE = LOGN < 5 ? LOGN : 5;
    typedef NttShape<LOGN, LOGE> S;
    static constexpr int CT = 256;
    static constexpr int SLOTS = CT / S::T;
    static constexpr int G = SLOTS / 8;                  // products per CTA
    static constexpr int PADN = Pad32<LOGN, LOGE>::padn;   // ntt_kernels.cuh: PadRule
    static constexpr int IPT = G * N / CT;               // positions per thread in the pointwise step
    static constexpr size_t SMEM = (size_t)G * 2 * N * 8 + (size_t)SLOTS * PADN * 4 + (size_t)G * 4 * N * 4;
    static_assert(S::T <= 32 && SLOTS % 8 == 0 && (G * N) % CT == 0, "unsupported ring degree for the fused Tn product");
};

struct TnParams {
    NttParams<Lazy32> P[2];
    Small32 ms[2];
    u64 mu[2];  // floor(2^64 / p_r)
    CrtParams cp;
};

__device__ __forceinline__ u32 tn_reduce64(u64 acc, u32 p, u64 mu) {  // acc mod p, canonical
    u64 r = acc - __umul64hi(acc, mu) * p;  // in [0, 2p)
    return (u32)(r >= p ? r - p : r);
}

template <int LOGN>
#ifndef FHE_TN_MINB
#define FHE_TN_MINB 3
#endif
__global__ void __launch_bounds__(256, FHE_TN_MINB)
tn_mul_fused_kernel(const __grid_constant__ TnParams X, const u64 *__restrict__ a, const u64 *__restrict__ b,
                    u64 *__restrict__ c, size_t batch) {
    typedef TnGeom<LOGN> G;
    typedef typename G::S S;
    constexpr int LOGE = G::LOGE, N = G::N, LAST = S::P - 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *xin = reinterpret_cast<u64 *>(smem_raw);                       // [G][2][N] operands
    u32 *xch = reinterpret_cast<u32 *>(xin + (size_t)G::G * 2 * N);     // [SLOTS][PADN]
    u32 *res1 = xch + (size_t)G::SLOTS * G::PADN;                       // [G][4][N] class residues mod p1
    const int t = threadIdx.x;
    const int slot = t / S::T, tid = t % S::T;
    u32 *sm = xch + (size_t)slot * G::PADN;
    const size_t g0 = (size_t)blockIdx.x * G::G;
    const int na = (int)(batch - g0 < (size_t)G::G ? batch - g0 : (size_t)G::G);

    for (int i = t; i < G::G * 2 * N; i += G::CT) {
        const int g = i / (2 * N), rem = i % (2 * N);
        u64 v = 0;
        if (g < na) v = rem < N ? a[(g0 + g) * N + rem] : b[(g0 + g) * N + rem - N];
        xin[i] = v;
    }
    __syncthreads();

#pragma unroll 1
    for (int r = 0; r < 2; r++) {
        const Small32 &ms = X.ms[r];
        const Lazy32 &ml = X.P[r].mod;
        {   // forward transform of limb (slot % 4) of operand (slot % 8) / 4 of product slot / 8
            const TwSrc<Small32> twf = {X.P[r].c_fwd, X.P[r].fwd};
            const u64 *xi = xin + (size_t)(slot >> 2) * N;  // [g][operand] rows are consecutive
            const int sh = 16 * (slot & 3);
            u32 x[S::E];
#pragma unroll
            for (int e = 0; e < S::E; e++) x[e] = (u32)(xi[S::pos(0, tid, e)] >> sh) & 0xffffu;
            fwd_chain<Small32, LOGN, LOGE>(x, sm, tid, ms, twf);
            // csub-free butterflies: x < 2^16 + 2*LOGN*p < 2^32; partial reduction below 2^28 (see extprod_fused.cu)
#pragma unroll
            for (int e = 0; e < S::E; e++) x[e] = x[e] - (x[e] >> 27) * ms.q;
            exch_put<Lazy32, LOGN, LOGE, LAST>(x, sm + Pad32<LOGN, LOGE>::idx(S::pos(LAST, tid, 0)));
        }
        __syncthreads();
        // limb convolution: class w = sum_{u+v=w} A_u * B_v (only w <= 3 reaches the low 64 bits of the product)
#pragma unroll
        for (int m = 0; m < G::IPT; m++) {
            const int item = t + G::CT * m, g = item >> LOGN, p = Pad32<LOGN, LOGE>::idx(item & (N - 1));
            u32 *base = xch + (size_t)g * 8 * G::PADN + p;
            u32 A[4], B[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                A[u] = base[(size_t)u * G::PADN];
                B[u] = base[(size_t)(4 + u) * G::PADN];
            }
            u32 cls[4];
#pragma unroll
            for (int w = 0; w < 4; w++) {
                u64 acc = 0;
#pragma unroll
                for (int u = 0; u <= w; u++) acc += (u64)A[u] * B[w - u];  // < 4 * 2^56
                cls[w] = tn_reduce64(acc, ml.q, X.mu[r]);
            }
            // this thread is the only reader of position p of these eight slots: overwrite the A slots in place
#pragma unroll
            for (int w = 0; w < 4; w++) base[(size_t)w * G::PADN] = cls[w];
        }
        __syncthreads();
        // inverse transforms of the class slots (slot % 8 < 4).  Warps span 32/T consecutive slots: for T >= 8 that
        // is at most 4 aligned slots (all live or all dead); narrower transforms run every warp, the lanes of dead
        // slots computing on their own scratch slot
        if (S::T < 8 || (slot & 7) < 4) {
            const TwSrc<Lazy32> twi = {X.P[r].c_inv, X.P[r].inv};
            u32 x[S::E];
            exch_get<Lazy32, LOGN, LOGE, LAST>(x, sm + Pad32<LOGN, LOGE>::idx(S::pos(LAST, tid, 0)));
            inv_chain<Lazy32, LOGN, LOGE, LAST>(x, sm, tid, ml, twi, X.P[r].ninv, X.P[r].s_ninv);
            if ((slot & 7) < 4) {
                // residues mod p1 go to res1, mod p2 to the (now dead) B-limb slot of the same product and limb
                u32 *R = r == 0 ? res1 + ((size_t)(slot >> 3) * 4 + (slot & 3)) * N : xch + (size_t)(slot + 4) * G::PADN;
#pragma unroll
                for (int e = 0; e < S::E; e++) R[S::pos(0, tid, e)] = ml.canon2(x[e]);
            }
        }
        __syncthreads();
    }
    // CRT lift and recombination: c = sum_w centre(CRT(r1_w, r2_w)) << 16w  (mod 2^64)
    for (int i = t; i < G::G * N; i += G::CT) {
        const int g = i >> LOGN, p = i & (N - 1);
        if (g >= na) continue;
        u64 acc = 0;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const u32 r1 = res1[((size_t)g * 4 + w) * N + p];
            const u32 r2 = xch[(size_t)(g * 8 + 4 + w) * G::PADN + p];
            acc += crt_centered(r1, r2, X.cp.p1, X.cp.p2, X.cp.p1_inv_mod_p2, X.cp.P, X.cp.halfP, X.cp.m2) << (16 * w);
        }
        c[(g0 + g) * N + p] = acc;
    }
}

template <int LOGN>
static int launch_tn_fused(const TorusCtx &tc, const u64 *a, const u64 *b, u64 *c, size_t batch, cudaStream_t st) {
    typedef TnGeom<LOGN> G;
    TnParams X;
    X.P[0] = tc.plan1->p32;
    X.P[1] = tc.plan2->p32;
    init_mod(X.ms[0], TORUS_P1);
    init_mod(X.ms[1], TORUS_P2);
    X.mu[0] = ~0ull / TORUS_P1;
    X.mu[1] = ~0ull / TORUS_P2;
    X.cp = tc.cp;
    const size_t grid = (batch + G::G - 1) / G::G;
    FHE_REQUIRE(grid <= 0x7fffffffull, "tn_mul: batch too large");
    static unsigned long long done_mask = 0;
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    auto kern = tn_mul_fused_kernel<LOGN>;
    if (!((done_mask >> (dev & 63)) & 1ull)) {
        FHE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
        done_mask |= 1ull << (dev & 63);
    }
    kern<<<(unsigned)grid, G::CT, G::SMEM, st>>>(X, a, b, c, batch);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

bool tn_mul_fused_supported(int logn) { return logn >= 6 && logn <= 10; }

int tn_mul_fused_device(const TorusCtx &tc, const u64 *a, const u64 *b, u64 *c, size_t batch, cudaStream_t st) {
    switch (tc.logn) {
        case 6: return launch_tn_fused<6>(tc, a, b, c, batch, st);
        case 7: return launch_tn_fused<7>(tc, a, b, c, batch, st);
        case 8: return launch_tn_fused<8>(tc, a, b, c, batch, st);
        case 9: return launch_tn_fused<9>(tc, a, b, c, batch, st);
        case 10: return launch_tn_fused<10>(tc, a, b, c, batch, st);
    }
    set_error("internal: fused Tn product called for an unsupported ring degree");
    return -1;
}

}  // namespace fhe
