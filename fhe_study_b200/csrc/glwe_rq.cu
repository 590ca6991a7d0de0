// glwe_rq.cu -- the gfhe layer over Rq: GLev<Rq> * Vec<Rq> (gfhe/src/glev.rs:67-80) and
// GLWE<Rq>::key_switch (gfhe/src/glwe.rs:126-137).  Both are gadget products: sum_r GLWE_r * v_r with every
// component an Rq product (ring_nq.rs:586-607).  The reference transforms every operand of every product
// (l*(k+1) forward + l*(k+1) inverse transforms per GLev product); Rq multiplication is exact in Z_q[X]/(X^n+1),
// so here the GLWE rows are transformed ONCE at load (resident handle), each digit polynomial is transformed
// once, the sum is taken in the NTT domain and only (k+1) inverse transforms are run per ciphertext.
#include <memory>

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
    *out = h.release();
    return 0;
}
void fhe_rq_glev_destroy(fhe_rq_glev *h) {
    if (!h) return;
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
