// glwe_rq.cu -- the gfhe layer over Rq: GLev<Rq> * Vec<Rq> (gfhe/src/glev.rs:67-80) and
// GLWE<Rq>::key_switch (gfhe/src/glwe.rs:126-137).  Both are gadget products: sum_r GLWE_r * v_r with every
// component an Rq product (ring_nq.rs:586-607).  The reference transforms every operand of every product
// (l*(k+1) forward + l*(k+1) inverse transforms per GLev product); Rq multiplication is exact in Z_q[X]/(X^n+1),
// so here the GLWE rows are transformed ONCE at load (resident handle), each digit polynomial is transformed
// once, the sum is taken in the NTT domain and only (k+1) inverse transforms are run per ciphertext.
#include <memory>

#include <stdlib.h>
#include <string.h>

#include "../../include/fhe_b200.h"
#include "plan.cuh"
#include "runtime.cuh"

namespace fhe {
int plan_launch(const fhe_ntt_plan *plan, int mode, const u64 *a, const u64 *b, u64 *c, u64 *c_evals, size_t batch,
                int flags, cudaStream_t st);  // lib_core.cu; mode 0 forward, 1 inverse
}
using namespace fhe;

struct fhe_rq_glev {
    fhe_ntt_plan *plan = nullptr;  // owned reference
    u64 k = 0, rows = 0;
    u64 *evals = nullptr;          // [rows][k+1][n] NTT images of the rows (reference order, canonical)
    u32 *evals_f = nullptr;        // fused-kernel layout [row][v][thread][4] (item = t + 256 (4v + j)), when instantiated
};

namespace {

constexpr int CH = 4;  // components accumulated per thread

// acc[b][c][x] = sum_r D[b][r][x] * E[r][c][x] mod q.  FAST: q < 2^32 and rows*(q-1)^2 < 2^64, plain u64 sums.
template <bool FAST>
__global__ void gadget_mac_kernel(const u64 *__restrict__ D, const u64 *__restrict__ E, u64 *__restrict__ acc, size_t batch,
                                  u32 n, u32 k1, u32 rows, u64 q) {
    const u32 chunks = (k1 + CH - 1) / CH;
    const size_t total = batch * (size_t)chunks * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 x = (u32)(idx % n);
        const u32 ch = (u32)((idx / n) % chunks);
        const size_t b = idx / ((size_t)n * chunks);
        const u32 c0 = ch * CH;
        u64 s[CH] = {0, 0, 0, 0};
        for (u32 r = 0; r < rows; r++) {
            const u64 d = D[(b * rows + r) * (size_t)n + x];
#pragma unroll
            for (int j = 0; j < CH; j++) {
                if (c0 + j < k1) {
                    const u64 e = E[((size_t)r * k1 + c0 + j) * n + x];
                    if (FAST) {
                        s[j] += d * e;
                    } else {
                        const u64 p = (u64)(((u128)d * (u128)e) % (u128)q);  // Zq::mul, zq.rs:315-328
                        const u64 v = s[j] + p;                              // Zq::add, zq.rs:219-231 (q < 2^63)
                        s[j] = v >= q ? v - q : v;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < CH; j++)
            if (c0 + j < k1) acc[(b * k1 + c0 + j) * (size_t)n + x] = FAST ? s[j] % q : s[j];
    }
}

// GLWE::key_switch's last line (glwe.rs:136): out = (0, .., 0, b) - rhs
__global__ void ks_finish_kernel(const u64 *__restrict__ ct, const u64 *__restrict__ rhs, u64 *__restrict__ out, size_t batch,
                                 u32 n, u32 k1, u64 q) {
    const size_t total = batch * (size_t)k1 * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 c = (u32)((idx / n) % k1);
        const u64 y = rhs[idx];
        const u64 x = c + 1 == k1 ? ct[idx] : 0;
        out[idx] = x >= y ? x - y : (q + x) - y;  // Zq::sub, zq.rs:259-277
    }
}

__global__ void rq_decompose_mask_kernel(const u64 *__restrict__ ct, u64 *__restrict__ out, size_t batch, u32 n, u32 k, u64 q,
                                         u32 beta, u32 l);

inline unsigned grid_for(size_t work, int threads = 256) {
    size_t g = (work + threads - 1) / threads;
    const size_t cap = (size_t)num_sms() * 16;
    return (unsigned)(g < 1 ? 1 : g > cap ? cap : g);
}

// out_b = sum_r rows_r * v_{b,r} for device-resident v ([batch][rows][n], canonical coefficients); out [batch][k+1][n]
int gadget_product_device(const fhe_rq_glev *h, const u64 *v, u64 *out, size_t batch, cudaStream_t st) {
    const u64 n = h->plan->host.n, q = h->plan->host.q, k1 = h->k + 1;
    int rc;
    Scratch sD, sA;  // NTT(v) and the NTT-domain sums (the transforms are never run in place)
    if ((rc = sD.alloc(batch * h->rows * n * 8, st))) return rc;
    if ((rc = sA.alloc(batch * k1 * n * 8, st))) return rc;
    u64 *v_scratch = sD.ptr<u64>(), *acc = sA.ptr<u64>();
    if ((rc = plan_launch(h->plan, 0 /* forward */, v, nullptr, v_scratch, nullptr, batch * h->rows, 0, st))) return rc;
    const bool fast = q < (1ull << 32) && (u128)h->rows * (u128)(q - 1) * (u128)(q - 1) < ((u128)1 << 64);
    const size_t work = batch * ((k1 + CH - 1) / CH) * n;
    if (fast) gadget_mac_kernel<true><<<grid_for(work), 256, 0, st>>>(v_scratch, h->evals, acc, batch, (u32)n, (u32)k1, (u32)h->rows, q);
    else gadget_mac_kernel<false><<<grid_for(work), 256, 0, st>>>(v_scratch, h->evals, acc, batch, (u32)n, (u32)k1, (u32)h->rows, q);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return plan_launch(h->plan, 1 /* inverse */, acc, nullptr, out, nullptr, batch * k1, 0, st);
}


// ---------------------------------------------------------------------------------------------------------------
// Fused GLWE<Rq>::key_switch for beta = 2 under the Small32 policy (q < 2^22, 2q*n <= 2^32: the reference's 65537):
// one CTA handles A ciphertexts; decompose -> digit NTTs -> MAC against the resident NTT images of the KSK rows ->
// k+1 inverse NTTs -> (0, b) - rhs, one HBM round trip.  Same structure as extprod_fused.cu with a single modulus:
// slots of T = n/32 threads run the digit transforms (csub-free butterflies, stage 0 on bits is a select), all 256
// threads then MAC the round's digits into 64-bit register accumulators (key loads shared by the A ciphertexts),
// finally the sums are reduced mod q and inverse-transformed.
// ---------------------------------------------------------------------------------------------------------------
template <int LOGN, int K1> struct KsGeom {
    static constexpr int N = 1 << LOGN;
    static constexpr int LOGE = LOGN < 5 ? LOGN : 5;
    typedef NttShape<LOGN, LOGE> S;
    static constexpr int CT = 256;
    static constexpr int SLOTS = CT / S::T;
    static constexpr int PADN = Pad32<LOGN, LOGE>::padn;   // ntt_kernels.cuh: PadRule
    static constexpr int ITEMS = K1 * N;                 // (component, position)
    static constexpr int IPT = (ITEMS + CT - 1) / CT;
    static constexpr int IPT4 = (IPT + 3) / 4 * 4;
    static constexpr int A = IPT <= 4 ? 4 : IPT <= 12 ? 2 : 1;   // ciphertexts per CTA (A * IPT 64-bit accumulators)
    static constexpr int DPR = SLOTS / A;                // digits per round
    static constexpr size_t SMEM = (size_t)A * (K1 - 1) * N * 4 + (size_t)SLOTS * PADN * 4;
    static_assert(S::T <= 32 && SLOTS % A == 0 && A * K1 <= SLOTS, "unsupported shape for the fused key switch");
};

struct KsParams {
    NttParams<Small32> P;
    u64 mu;            // floor(2^64 / q)
    const u32 *Rf;     // fused key layout
    u32 l, nd;         // levels, digits per ciphertext = k*l
};

template <int LOGN, int K1>
__global__ void __launch_bounds__(256, 2)
glwe_ks_fused_kernel(const __grid_constant__ KsParams X, const u64 *__restrict__ ct, u64 *__restrict__ out, size_t batch) {
    typedef KsGeom<LOGN, K1> G;
    typedef typename G::S S;
    constexpr int LOGE = G::LOGE, N = G::N, LAST = S::P - 1, A = G::A, K = K1 - 1, GLWE = K1 * N;
    constexpr int G0 = 1 << S::g(0), H0 = G0 >> 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32 *xin = reinterpret_cast<u32 *>(smem_raw);              // [A][K][N] mask coefficients (canonical, < q < 2^22)
    u32 *xch = xin + (size_t)A * K * N;                        // [SLOTS][PADN]
    const Small32 &ms = X.P.mod;
    const int t = threadIdx.x;
    const int slot = t / S::T, tid = t % S::T;
    const int s_acc = slot % A, s_dig = slot / A;
    u32 *sm = xch + (size_t)slot * G::PADN;
    const size_t acc0 = (size_t)blockIdx.x * A;
    const int na = (int)(batch - acc0 < (size_t)A ? batch - acc0 : (size_t)A);
    const u32 l = X.l;
    const u32 sat_thr = l >= 32 ? 0xffffffffu : (1u << l);     // Zq::decompose: v >= 2^l -> every digit is 1 (zq.rs:174-186)

    for (int i = t; i < A * K * N; i += G::CT) {
        const int aa = i / (K * N), rem = i % (K * N);
        xin[i] = aa < na ? (u32)ct[(acc0 + aa) * GLWE + rem] : 0u;
    }
    __syncthreads();

    u64 acc[A][G::IPT];
#pragma unroll
    for (int aa = 0; aa < A; aa++)
#pragma unroll
        for (int m = 0; m < G::IPT; m++) acc[aa][m] = 0;
    const TwSrc<Small32> twf = {X.P.c_fwd, X.P.fwd};
    const int rounds = (int)((X.nd + G::DPR - 1) / G::DPR);
#pragma unroll 1
    for (int round = 0; round < rounds; round++) {
        const u32 d = (u32)round * G::DPR + s_dig;
        {   // every slot runs the transform (the exchanges inside synchronise whole warps, and k*l need not fill the
            // last round): slots past the last digit transform zeros into their own slot, which the MAC never reads
            const bool live = d < X.nd;
            const u32 dc = live ? d : 0;
            const u32 *xi = xin + ((size_t)s_acc * K + dc / l) * N;
            const u32 sh = l - 1 - dc % l;
            u32 x[S::E];
#pragma unroll
            for (int e = 0; e < S::E; e++) {
                const u32 v = xi[S::pos(0, tid, e)];
                const u32 bit = (l < 32 && v >= sat_thr) ? 1u : (sh < 32 ? (v >> sh) & 1u : 0u);
                x[e] = live ? bit : 0u;
            }
            {   // stage 0 on bits: V = b*S is a select
                const u32 S1 = twf.c0[1].w;
#pragma unroll
                for (int qi = 0; qi < (S::E >> S::g(0)); qi++)
#pragma unroll
                    for (int lo = 0; lo < H0; lo++) {
                        const u32 U = x[qi * G0 + lo], V = (0u - x[qi * G0 + lo + H0]) & S1;
                        x[qi * G0 + lo] = U + V;
                        x[qi * G0 + lo + H0] = U + ms.q2 - V;
                    }
            }
            fwd_pass<Small32, LOGN, LOGE, 0, 1>(x, tid, ms, twf);
            if constexpr (S::P > 1) fwd_chain<Small32, LOGN, LOGE, 1>(x, sm, tid, ms, twf);
            exch_put<Lazy32, LOGN, LOGE, LAST>(x, sm + Pad32<LOGN, LOGE>::idx(S::pos(LAST, tid, 0)));  // < (2 LOGN + 1) q: no reduction needed
        }
        __syncthreads();
        const int nd = min(G::DPR, (int)X.nd - round * G::DPR);
        const uint4 *Rt = reinterpret_cast<const uint4 *>(X.Rf) + (size_t)round * G::DPR * (G::IPT4 / 4) * G::CT + t;
#pragma unroll 2
        for (int dd = 0; dd < nd; dd++) {
            u32 rv[G::IPT4];
#pragma unroll
            for (int v = 0; v < G::IPT4 / 4; v++) {
                const uint4 q4 = __ldg(Rt + (size_t)(dd * (G::IPT4 / 4) + v) * G::CT);
                rv[4 * v] = q4.x; rv[4 * v + 1] = q4.y; rv[4 * v + 2] = q4.z; rv[4 * v + 3] = q4.w;
            }
#pragma unroll
            for (int aa = 0; aa < A; aa++) {
                const u32 *D = xch + (size_t)(dd * A + aa) * G::PADN;
#pragma unroll
                for (int m = 0; m < G::IPT; m++) {
                    const int item = t + G::CT * m;
                    acc[aa][m] += (u64)D[Pad32<LOGN, LOGE>::idx(item & (N - 1))] * rv[m];  // key padding beyond ITEMS is zero
                }
            }
        }
        __syncthreads();
    }
    // sums -> inverse-transform inputs (slot aa*K1 + c)
#pragma unroll
    for (int aa = 0; aa < A; aa++)
#pragma unroll
        for (int m = 0; m < G::IPT; m++) {
            const int item = t + G::CT * m;
            if (item < G::ITEMS) {
                const u64 a = acc[aa][m];
                u64 r = a - __umul64hi(a, X.mu) * ms.q;
                if (r >= ms.q) r -= ms.q;
                xch[(size_t)(aa * K1 + (item >> LOGN)) * G::PADN + Pad32<LOGN, LOGE>::idx(item & (N - 1))] = (u32)r;
            }
        }
    __syncthreads();
    if ((t & ~31) / S::T < A * K1) {  // whole warps run the transform (see extprod_fused.cu)
        const TwSrc<Small32> twi = {X.P.c_inv, X.P.inv};
        u32 x[S::E];
        exch_get<Lazy32, LOGN, LOGE, LAST>(x, sm + Pad32<LOGN, LOGE>::idx(S::pos(LAST, tid, 0)));
        inv_chain<Small32, LOGN, LOGE, LAST>(x, sm, tid, ms, twi, X.P.ninv, X.P.s_ninv);
        const int aa = slot / K1, c = slot % K1;
        if (slot < A * K1 && aa < na) {
            // GLWE::key_switch's last line (glwe.rs:136): (0, .., 0, b) - rhs, Zq::sub (zq.rs:259-277)
            const size_t row = (acc0 + aa) * GLWE + (size_t)c * N;
#pragma unroll
            for (int e = 0; e < S::E; e++) {
                const int p = S::pos(0, tid, e);
                const u64 y = ms.canon2(x[e]);
                const u64 xv = c == K ? ct[row + p] : 0;
                out[row + p] = xv >= y ? xv - y : (ms.q + xv) - y;
            }
        }
    }
}

// evals [row][c][x] (u64) -> fused layout (u32 [row][v][t][4], item = c*N + x = t + 256 (4v + j)), zero padded
__global__ void glev_fused_layout_kernel(const u64 *__restrict__ E, u32 *__restrict__ Ef, size_t rows, int items, int ipt4) {
    const size_t total = rows * 256 * (size_t)ipt4;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx & 3), t = (int)((idx >> 2) & 255), v = (int)((idx >> 10) % (ipt4 / 4));
        const size_t r = idx / ((size_t)ipt4 * 256);
        const int item = t + 256 * (4 * v + j);
        Ef[idx] = item < items ? (u32)E[r * items + item] : 0u;
    }
}

#define FHE_KS_SHAPES(F) F(7, 17) F(10, 2) F(9, 2) F(8, 2) F(6, 2) F(6, 5) F(7, 2) F(8, 3)

bool ks_fused_supported(const fhe_ntt_plan *plan, u64 k, u64 rows) {
    if (plan->kind != 3) return false;  // Small32 only
    if (plan->loge != (plan->logn < 5 ? plan->logn : 5)) return false;  // table order of 32 coefficients per thread
    const u64 q = plan->host.q;
    // exactness of the 64-bit sums: rows * (2 logn + 1) q * q < 2^64
    if ((u128)rows * (2 * plan->logn + 1) * q * q >= ((u128)1 << 64)) return false;
#define F(L, K) if (plan->logn == L && (int)k + 1 == K) return true;
    FHE_KS_SHAPES(F)
#undef F
    return false;
}
int ks_fused_ipt4(int logn, int k1) {
    const int ipt = (k1 * (1 << logn) + 255) / 256;
    return (ipt + 3) / 4 * 4;
}

template <int LOGN, int K1>
int launch_ks_fused(const fhe_rq_glev *h, u32 l, const u64 *ct, u64 *out, size_t batch, cudaStream_t st) {
    typedef KsGeom<LOGN, K1> G;
    KsParams X;
    X.P = h->plan->psm;
    X.mu = ~0ull / h->plan->host.q;
    X.Rf = h->evals_f;
    X.l = l;
    X.nd = (u32)h->rows;
    auto kern = glwe_ks_fused_kernel<LOGN, K1>;
    static unsigned long long done_mask = 0;
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    if (!((done_mask >> (dev & 63)) & 1ull)) {
        FHE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));
        done_mask |= 1ull << (dev & 63);
    }
    const size_t grid = (batch + G::A - 1) / G::A;
    FHE_REQUIRE(grid <= 0x7fffffffull, "key switch: batch too large");
    kern<<<(unsigned)grid, G::CT, G::SMEM, st>>>(X, ct, out, batch);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}
int ks_fused_device(const fhe_rq_glev *h, u32 l, const u64 *ct, u64 *out, size_t batch, cudaStream_t st) {
    const int logn = h->plan->logn, k1 = (int)h->k + 1;
#define F(L, K) if (logn == L && k1 == K) return launch_ks_fused<L, K>(h, l, ct, out, batch, st);
    FHE_KS_SHAPES(F)
#undef F
    set_error("internal: fused key switch called for an unsupported shape");
    return -1;
}
}  // namespace

extern "C" {

int fhe_rq_glev_load(const fhe_ntt_plan *plan, uint64_t k, uint64_t rows, const uint64_t *glwes, fhe_rq_glev **out) {
    FHE_REQUIRE(out != nullptr && glwes != nullptr && plan != nullptr, "fhe_rq_glev_load: null pointer");
    *out = nullptr;
    FHE_REQUIRE(k >= 1 && rows >= 1 && rows <= (1u << 20), "fhe_rq_glev_load: need k >= 1 and 1 <= rows <= 2^20");
    std::unique_ptr<fhe_rq_glev> h(new fhe_rq_glev());
    int rc = fhe_ntt_plan_create(plan->host.q, plan->host.n, &h->plan);  // take a reference on the cached plan
    if (rc) return rc;
    h->k = k;
    h->rows = rows;
    const size_t polys = rows * (k + 1), bytes = polys * plan->host.n * sizeof(u64);
    cudaStream_t st = current_stream();
    IoBuf br;
    cudaError_t e = cudaMalloc((void **)&h->evals, bytes);
    if (e == cudaSuccess) rc = br.init(glwes, bytes, true, false, st);
    if (e == cudaSuccess && !rc) rc = plan_launch(h->plan, 0, br.ptr<u64>(), nullptr, h->evals, nullptr, polys, 0, st);
    if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess || rc) {
        if (e != cudaSuccess) set_error(std::string("fhe_rq_glev_load: ") + cudaGetErrorString(e));
        cudaFree(h->evals);
        fhe_ntt_plan_destroy(h->plan);
        return rc ? rc : -2;
    }
    if (ks_fused_supported(h->plan, k, rows)) {  // fused key-switch layout (used when the call asks for beta = 2)
        const int items = (int)((k + 1) * plan->host.n), ipt4 = ks_fused_ipt4(h->plan->logn, (int)k + 1);
        const size_t words = rows * 256 * (size_t)ipt4;
        e = cudaMalloc((void **)&h->evals_f, words * sizeof(u32));
        if (e == cudaSuccess) {
            glev_fused_layout_kernel<<<grid_for(words), 256, 0, st>>>(h->evals, h->evals_f, rows, items, ipt4);
            count_launch(1);
            e = cudaStreamSynchronize(st);
        }
        if (e != cudaSuccess) {
            set_error(std::string("fhe_rq_glev_load: ") + cudaGetErrorString(e));
            cudaFree(h->evals_f);
            cudaFree(h->evals);
            fhe_ntt_plan_destroy(h->plan);
            return -2;
        }
    }
    *out = h.release();
    return 0;
}
void fhe_rq_glev_destroy(fhe_rq_glev *h) {
    if (!h) return;
    cudaFree(h->evals_f);
    cudaFree(h->evals);
    fhe_ntt_plan_destroy(h->plan);
    delete h;
}

int fhe_rq_glev_mul(const fhe_rq_glev *h, const uint64_t *v, uint64_t *out, size_t batch) {
    FHE_REQUIRE(h != nullptr, "null GLev handle");
    if (batch == 0) return 0;
    FHE_REQUIRE(v && out, "fhe_rq_glev_mul: null pointer");
    cudaStream_t st = current_stream();
    const u64 n = h->plan->host.n;
    const size_t vbytes = batch * h->rows * n * 8;
    IoBuf bv, bo;
    int rc;
    if ((rc = bv.init(v, vbytes, true, false, st))) return rc;
    if ((rc = bo.init(out, batch * (h->k + 1) * n * 8, false, true, st))) return rc;
    if ((rc = gadget_product_device(h, bv.ptr<u64>(), bo.ptr<u64>(), batch, st))) return rc;
    return finish_all({&bv, &bo}, st);
}

int fhe_glwe_rq_key_switch(const fhe_rq_glev *ksk, uint32_t beta, uint32_t l, const uint64_t *ct, uint64_t *out, size_t batch) {
    FHE_REQUIRE(ksk != nullptr, "null KSK handle");
    if (batch == 0) return 0;
    FHE_REQUIRE(ct && out, "fhe_glwe_rq_key_switch: null pointer");
    FHE_REQUIRE(beta >= 2 && l >= 1 && ksk->rows == ksk->k * (u64)l, "fhe_glwe_rq_key_switch: the handle must hold k*l rows");
    const u64 n = ksk->plan->host.n, q = ksk->plan->host.q, k = ksk->k, k1 = k + 1;
    if (beta != 2)
        for (uint32_t i = 1, pw = 1; i <= l; i++) {
            pw *= beta;
            FHE_REQUIRE(pw != 0 && q / pw != 0, "fhe_glwe_rq_key_switch: q / beta^i is zero (the reference divides by zero here)");
        }
    cudaStream_t st = current_stream();
    IoBuf bi, bo;
    Scratch sd, sr;
    int rc;
    if ((rc = bi.init(ct, batch * k1 * n * 8, true, false, st))) return rc;
    if ((rc = bo.init(out, batch * k1 * n * 8, false, true, st))) return rc;
    {   // fused kernel (beta = 2, shapes instantiated, Small32 modulus); FHE_GLWE_KS_PATH=unfused forces the blocks below
        const char *force = getenv("FHE_GLWE_KS_PATH");
        if (ksk->evals_f != nullptr && beta == 2 && l <= 63 && bi.ptr<u64>() != bo.ptr<u64>() &&
            !(force && strcmp(force, "unfused") == 0)) {
            if ((rc = ks_fused_device(ksk, l, bi.ptr<u64>(), bo.ptr<u64>(), batch, st))) return rc;
            return finish_all({&bi, &bo}, st);
        }
    }
    if ((rc = sd.alloc(batch * k * l * n * 8, st))) return rc;
    if ((rc = sr.alloc(batch * k1 * n * 8, st))) return rc;
    // digits of the mask polynomials: [b][i][j][n]  (a_i.decompose(beta, l), ring_nq.rs:67-77); the body is skipped
    rq_decompose_mask_kernel<<<grid_for(batch * k * n), 256, 0, st>>>(bi.ptr<u64>(), sd.ptr<u64>(), batch, (u32)n, (u32)k, q, beta, l);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    if ((rc = gadget_product_device(ksk, sd.ptr<u64>(), sr.ptr<u64>(), batch, st))) return rc;
    ks_finish_kernel<<<grid_for(batch * k1 * n), 256, 0, st>>>(bi.ptr<u64>(), sr.ptr<u64>(), bo.ptr<u64>(), batch, (u32)n, (u32)k1, q);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&bi, &bo}, st);
}

}  // extern "C"

namespace {
__device__ __forceinline__ u64 zq_from_u64_(u64 q, u64 v) { return v >= q ? v % q : v; }
__device__ __forceinline__ u32 pow_u32_wrapping_(u32 b, u32 e) {
    u32 r = 1;
    for (u32 i = 0; i < e; i++) r *= b;
    return r;
}
// Zq::decompose (zq.rs:140-186) of every mask coefficient of `batch` GLWEs: out[((b*k + i)*l + j)*n + c]
__global__ void rq_decompose_mask_kernel(const u64 *__restrict__ ct, u64 *__restrict__ out, size_t batch, u32 n, u32 k, u64 q,
                                         u32 beta, u32 l) {
    const size_t total = batch * (size_t)k * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 c = (u32)(idx % n);
        const size_t poly = idx / n;             // b*k + i
        const size_t b = poly / k;
        const u32 i = (u32)(poly % k);
        const u64 v = ct[(b * (k + 1) + i) * (size_t)n + c];
        u64 *o = out + poly * (size_t)l * n + c;
        if (beta == 2) {  // zq.rs:174-186
            const bool sat = v >= ((u64)1 << (l & 63));
            for (u32 j = 0; j < l; j++) {
                const u32 sh = l - 1 - j;
                o[(size_t)j * n] = sat ? 1 : zq_from_u64_(q, sh < 64 ? ((v >> sh) & 1) : 0);
            }
        } else {  // zq.rs:147-172
            u64 rem = v;
            const bool sat = rem >= (u64)pow_u32_wrapping_(beta, l);
            for (u32 lv = 1; lv <= l; lv++) {
                if (sat) { o[(size_t)(lv - 1) * n] = (u64)beta - 1; continue; }
                const u64 den = q / (u64)pow_u32_wrapping_(beta, lv);
                const u64 x_i = rem / den;
                o[(size_t)(lv - 1) * n] = zq_from_u64_(q, x_i);
                if (x_i != 0) rem = rem % den;
            }
        }
    }
}
}  // namespace
