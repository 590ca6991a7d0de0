// torus_kernels.cu -- exact arithmetic in T_q[X]/(X^N+1), q = 2^64 (the reference's Tn, arith/src/ring_torus.rs)
// and the TFHE external product / CMux built on it (tfhe/src/tggsw.rs:39-62).
//
// The reference multiplies torus polynomials with an O(N^2) u128 schoolbook, negacyclic fold by
// wrapping_sub and truncation to u64 (ring_torus.rs:266-298): that IS the exact negacyclic convolution
// in Z_{2^64}[X]/(X^N+1), so any exact algorithm is bit-identical.  Here the convolution is
// computed over the integers with two 27-bit NTT primes (P = p1*p2 ~ 2^54, torus.cuh) on small limbs, lifted to the
// centred representative by CRT, and recombined mod 2^64:
//   * Tn*Tn     : both operands in four 16-bit limbs; plane products are bounded by N*2^32, the four
//                 weight classes 2^(16w), w = 0..3, by 4*N*2^32 < P/2 (N <= 2^15);
//   * ext. prod.: the TGLWE input is decomposed into its 64 bit-planes (Tn::decompose, beta=2,l=64:
//                 torus.rs:43-52), the TGGSW rows into two 32-bit limbs transformed ONCE at load time;
//                 sum over the (k+1)*64 rows is bounded by (k+1)*64*N*2^32 < P/2 (checked at load).
// This file holds the unfused building blocks (split / MAC / CRT kernels) around the batched NTT kernels.
#include "../../include/fhe_b200.h"
#include "ntt_kernels.cuh"
#include "plan_host.hpp"
#include "runtime.cuh"
#include "torus.cuh"

#include <vector>

#include <stdlib.h>

namespace fhe {


// ---------------------------------------------------------------------------------------------------
// elementwise torus kernels
// ---------------------------------------------------------------------------------------------------
// planes[(poly*4 + w)*n + c] = 16-bit limb w of a[poly*n + c]
__global__ void split16_kernel(const u64 *__restrict__ a, u64 *__restrict__ planes, size_t polys, u32 n) {
    const size_t total = polys * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t poly = i / n, c = i % n;
        const u64 v = a[i];
#pragma unroll
        for (int w = 0; w < 4; w++) planes[(poly * 4 + w) * n + c] = (v >> (16 * w)) & 0xffffull;
    }
}
// Tn::decompose(2,64) (ring_torus.rs:67-77 + torus.rs:43-52): plane j holds bit (63-j) of every coefficient.
// planes[(poly*64 + j)*n + c]
__global__ void bitplanes_kernel(const u64 *__restrict__ a, u64 *__restrict__ planes, size_t polys, u32 n) {
    const size_t total = polys * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t poly = i / n, c = i % n;
        const u64 v = a[i];
#pragma unroll 8
        for (int j = 0; j < 64; j++) planes[(poly * 64 + j) * n + c] = (v >> (63 - j)) & 1ull;
    }
}
// rows: u64 values -> limb planes reduced mod p: out[(row*2 + limb)*n + c] = ((v >> 32*limb) & 0xffffffff) % p
__global__ void split32_mod_kernel(const u64 *__restrict__ rows, u64 *__restrict__ out, size_t polys, u32 n, u32 p) {
    const size_t total = polys * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t poly = i / n, c = i % n;
        const u64 v = rows[i];
        out[(poly * 2 + 0) * n + c] = (u32)v % p;
        out[(poly * 2 + 1) * n + c] = (u32)(v >> 32) % p;
    }
}

// weight-class products of the limb planes in the NTT domain (Tn*Tn):
//   C[(poly*4 + w)*n + x] = sum_{u+v=w} A[(poly*4+u)*n + x] * B[(poly*4+v)*n + x]  mod p
__global__ void limb_conv_kernel(const u64 *__restrict__ A, const u64 *__restrict__ B, u64 *__restrict__ C,
                                 size_t polys, u32 n, Lazy32 m) {
    const size_t total = polys * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t poly = i / n, x = i % n;
        u32 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            a[u] = (u32)A[(poly * 4 + u) * n + x];
            b[u] = (u32)B[(poly * 4 + u) * n + x];
        }
#pragma unroll
        for (int w = 0; w < 4; w++) {
            u32 acc = 0;
#pragma unroll
            for (int u = 0; u <= w; u++) acc = m.csub(acc + m.mul(a[u], b[w - u]), m.q);
            C[(poly * 4 + w) * n + x] = acc;
        }
    }
}

// External-product MAC in the NTT domain (tggsw.rs:45-62 with tggsw.rs:139-149):
//   out[((b*(k+1) + c)*2 + limb)*n + x] = sum_{d < (k+1)*64} D[(b*(k+1)*64 + d)*n + x] * R[((d*(k+1) + c)*2 + limb)*n + x]
// D = NTT of the bit-planes of accumulator b, R = NTT of the limb planes of the TGGSW rows.
__global__ void extprod_mac_kernel(const u64 *__restrict__ D, const u64 *__restrict__ R, u64 *__restrict__ out,
                                   size_t batch, u32 n, u32 k1, Lazy32 m) {
    const u32 items = k1 * 2 * n;  // (c, limb, x)
    const size_t total = batch * items;
    const u32 nd = k1 * 64;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / items;
        const u32 item = (u32)(i % items), x = item % n;
        const u64 *Db = D + b * nd * (size_t)n + x;
        const u64 *Rp = R + item;
        u64 acc = 0;  // products < 2^60: fold every 8 terms
        for (u32 d = 0; d < nd; d += 8) {
#pragma unroll
            for (u32 dd = 0; dd < 8; dd++) {
                const u32 dv = (u32)Db[(size_t)(d + dd) * n];
                const u32 rv = (u32)Rp[(size_t)(d + dd) * items];
                acc += (u64)m.mul(dv, rv);
            }
        }
        out[i] = acc % m.q;
    }
}

// CRT lift + recombination: res1/res2 hold residues mod p1/p2 of W integer polynomials per output
// polynomial ([poly][w][n]); out[poly*n + c] = sum_w centre(CRT(r1, r2)) << shift[w]   (mod 2^64),
// optionally + addend[poly*n + c] (the CMux's ct1, tggsw.rs:39-41).
__global__ void crt_recombine_kernel(const u64 *__restrict__ res1, const u64 *__restrict__ res2,
                                     const u64 *__restrict__ addend, u64 *__restrict__ out, size_t polys, u32 n,
                                     CrtParams cp) {
    const size_t total = polys * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t poly = i / n, c = i % n;
        u64 acc = addend ? addend[i] : 0;
        for (int w = 0; w < cp.W; w++) {
            const u32 r1 = (u32)res1[(poly * cp.W + w) * n + c];
            const u32 r2 = (u32)res2[(poly * cp.W + w) * n + c];
            acc += crt_centered(r1, r2, cp.p1, cp.p2, cp.p1_inv_mod_p2, cp.P, cp.halfP, cp.m2) << cp.shift[w];
        }
        out[i] = acc;
    }
}

// wrapping elementwise ops on Tn / T64 vectors (ring_torus.rs:153-249, torus.rs:80-153): op 0 add, 1 sub, 2 neg
__global__ void tn_addsub_kernel(const u64 *a, const u64 *b, u64 *c, size_t len,
                                 int op) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x)
        c[i] = op == 0 ? a[i] + b[i] : op == 1 ? a[i] - b[i] : (u64)0 - a[i];
}
// Tn::left_rotate(h) (ring_torus.rs:118-132): out = [c_h..c_{n-1}, -c_0..-c_{h-1}], h reduced mod n.
// One rotation amount per polynomial group: h = hs[poly / group] (hs == nullptr: h_const).
// hs_stride: distance between consecutive groups' amounts in hs.  negacyclic != 0 (extension, no reference
// counterpart): h is taken mod 2n and h >= n negates, i.e. the true multiplication by X^{-h}.
__global__ void tn_left_rotate_kernel(const u64 *__restrict__ a, u64 *__restrict__ out, size_t polys, u32 n,
                                      const u64 *__restrict__ hs, u64 h_const, u32 group, size_t hs_stride,
                                      int negacyclic) {
    const size_t total = polys * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t poly = i / n;
        const u32 c = (u32)(i % n);
        const u64 hraw = hs ? hs[(poly / group) * hs_stride] : h_const;
        const u32 h = (u32)(hraw % n);
        const u32 src = c + h;
        const u64 v = src < n ? a[poly * n + src] : (u64)0 - a[poly * n + (src - n)];
        out[i] = (negacyclic && ((hraw / n) & 1)) ? (u64)0 - v : v;
    }
}

static inline unsigned grid_for(size_t work, int threads = 256) {
    size_t g = (work + threads - 1) / threads;
    const size_t cap = (size_t)num_sms() * 16;
    return (unsigned)(g < 1 ? 1 : g > cap ? cap : g);
}

// ---------------------------------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------------------------------
int TorusCtx::init(u64 n) {
    this->n = n;
    FHE_REQUIRE(n >= 2 && (n & (n - 1)) == 0 && n <= (1u << 15), "torus ring degree must be a power of two in 2..2^15");
    int rc;
    if ((rc = fhe_ntt_plan_create(TORUS_P1, n, &plan1))) return rc;
    if ((rc = fhe_ntt_plan_create(TORUS_P2, n, &plan2))) return rc;
    logn = hp_ilog2(n);
    init_mod(m1, TORUS_P1);
    init_mod(m2, TORUS_P2);
    cp.p1 = (u32)TORUS_P1;
    cp.p2 = (u32)TORUS_P2;
    cp.p1_inv_mod_p2 = (u32)hp_powmod(TORUS_P1 % TORUS_P2, TORUS_P2 - 2, TORUS_P2);
    cp.P = TORUS_P1 * TORUS_P2;
    cp.halfP = cp.P / 2;
    cp.m2 = m2;
    return 0;
}
TorusCtx::~TorusCtx() {
    fhe_ntt_plan_destroy(plan1);
    fhe_ntt_plan_destroy(plan2);
}

// forward / inverse NTT of `polys` device polynomials under prime index r (0: p1, 1: p2)
int TorusCtx::ntt(int r, int mode, const u64 *in, u64 *out, size_t polys, cudaStream_t st) const {
    const NttParams<Lazy32> &P = *reinterpret_cast<const NttParams<Lazy32> *>(plan_params32(r == 0 ? plan1 : plan2));
    int rc = ntt_launch_lazy32(logn, (r == 0 ? plan1 : plan2)->loge, mode, P, in, nullptr, out, nullptr, polys, 0, st);
    if (!rc) count_launch(1);
    return rc;
}

// c = a * b in Tn, `batch` independent products; device pointers.
// Tn * Tn for `batch` products; device pointers.  64 <= n <= 1024 runs the fused kernel (tn_fused.cu);
// FHE_TN_PATH=unfused forces the building-block path below (the tests cover both).
int tn_mul_device(const TorusCtx &tc, const u64 *a, const u64 *b, u64 *c, size_t batch, cudaStream_t st) {
    const char *force = getenv("FHE_TN_PATH");
    if (tn_mul_fused_supported(tc.logn) && tc.plan1->loge == (tc.logn < 5 ? tc.logn : 5) && tc.plan2->loge == tc.plan1->loge &&
        !(force && strcmp(force, "unfused") == 0))
        return tn_mul_fused_device(tc, a, b, c, batch, st);
    const u32 n = (u32)tc.n;
    const size_t plane_words = batch * 4 * (size_t)n;
    Scratch scratch;  // A planes, B planes, C residues for p1 and p2
    int rc = scratch.alloc(plane_words * 4 * sizeof(u64), st);
    if (rc) return rc;
    u64 *buf = scratch.ptr<u64>();
    u64 *A = buf, *B = buf + plane_words, *C1 = B + plane_words, *C2 = C1 + plane_words;
    for (int r = 0; r < 2 && !rc; r++) {
        split16_kernel<<<grid_for(batch * n), 256, 0, st>>>(a, A, batch, n);
        split16_kernel<<<grid_for(batch * n), 256, 0, st>>>(b, B, batch, n);
        count_launch(2);
        if ((rc = tc.ntt(r, MODE_FWD, A, A, batch * 4, st))) break;
        if ((rc = tc.ntt(r, MODE_FWD, B, B, batch * 4, st))) break;
        u64 *C = r == 0 ? C1 : C2;
        limb_conv_kernel<<<grid_for(batch * n), 256, 0, st>>>(A, B, C, batch, n, r == 0 ? tc.m1 : tc.m2);
        count_launch(1);
        rc = tc.ntt(r, MODE_INV, C, C, batch * 4, st);
    }
    if (!rc) {
        CrtParams cp = tc.cp;
        cp.W = 4;
        for (int w = 0; w < 4; w++) cp.shift[w] = 16 * w;
        crt_recombine_kernel<<<grid_for(batch * n), 256, 0, st>>>(C1, C2, nullptr, c, batch, n, cp);
        count_launch(1);
        if (cudaGetLastError() != cudaSuccess) { set_error("tn_mul kernel launch failed"); rc = -2; }
    }
    return rc;
}

// out = tggsw (x) ct1, or ct1 + tggsw (x) (ct2 - ct1) when ct2 != nullptr (TGGSW::cmux, tggsw.rs:39-41);
// `batch` TGLWE accumulators sharing one TGGSW; device pointers.  Dispatches to the fused kernel when the
// shape is instantiated (FHE_EXTPROD_PATH=fused|unfused forces a path; the tests cover both).
int extprod_device(const Tggsw &g, const u64 *ct1, const u64 *ct2, u64 *out, size_t batch, cudaStream_t st) {
    const char *force = getenv("FHE_EXTPROD_PATH");
    if (force && strncmp(force, "fused", 5) == 0 && g.R1f == nullptr) {
        set_error("FHE_EXTPROD_PATH=fused but this (n, k) has no fused instantiation");
        return -1;
    }
    if (g.R1f != nullptr && !(force && strcmp(force, "unfused") == 0))
        return extprod_fused_device(g, ct1, ct2, out, batch, ct2 != nullptr, st);

    const TorusCtx &tc = *g.tc;
    const u32 n = (u32)tc.n, k1 = (u32)g.k + 1;
    const size_t nd = (size_t)k1 * 64;                 // digit polynomials per accumulator
    const size_t chunk_max = std::max<size_t>(1, (512ull << 20) / (nd * n * sizeof(u64)));  // <= 512 MiB of planes
    const size_t chunk = std::min(batch, chunk_max);
    Scratch s_planes, s_D, s_res, s_diff;
    int rc;
    if ((rc = s_planes.alloc(chunk * nd * n * sizeof(u64), st))) return rc;
    if ((rc = s_D.alloc(chunk * nd * n * sizeof(u64), st))) return rc;
    if ((rc = s_res.alloc(2 * chunk * k1 * 2 * n * sizeof(u64), st))) return rc;
    if (ct2 && (rc = s_diff.alloc(chunk * k1 * n * sizeof(u64), st))) return rc;
    u64 *planes = s_planes.ptr<u64>(), *D = s_D.ptr<u64>(), *res = s_res.ptr<u64>(), *diff = s_diff.ptr<u64>();
    for (size_t b0 = 0; b0 < batch && !rc; b0 += chunk) {
        const size_t nb = std::min(chunk, batch - b0);
        const size_t res_words = nb * k1 * 2 * n;
        const u64 *in = ct1 + b0 * k1 * n;
        if (ct2) {
            if ((rc = tn_addsub_device(ct2 + b0 * k1 * n, ct1 + b0 * k1 * n, diff, nb * k1 * n, 1, st))) break;
            in = diff;
        }
        bitplanes_kernel<<<grid_for(nb * k1 * n), 256, 0, st>>>(in, planes, nb * k1, n);
        count_launch(1);
        for (int r = 0; r < 2 && !rc; r++) {
            if ((rc = tc.ntt(r, MODE_FWD, planes, D, nb * nd, st))) break;
            u64 *R = res + (size_t)r * chunk * k1 * 2 * n;
            extprod_mac_kernel<<<grid_for(res_words), 256, 0, st>>>(D, r == 0 ? g.R1 : g.R2, R, nb, n, k1,
                                                                  r == 0 ? tc.m1 : tc.m2);
            count_launch(1);
            rc = tc.ntt(r, MODE_INV, R, R, nb * k1 * 2, st);
        }
        if (rc) break;
        CrtParams cp = tc.cp;
        cp.W = 2;
        cp.shift[0] = 0;
        cp.shift[1] = 32;
        crt_recombine_kernel<<<grid_for(nb * k1 * n), 256, 0, st>>>(res, res + chunk * k1 * 2 * n,
                                                                   ct2 ? ct1 + b0 * k1 * n : nullptr,
                                                                   out + b0 * k1 * n, nb * k1, n, cp);
        count_launch(1);
        if (cudaGetLastError() != cudaSuccess) { set_error("extprod kernel launch failed"); rc = -2; }
    }
    return rc;
}

// Transforms the TGGSW rows once: R_r[((d*(k+1) + c)*2 + limb)*n + x] = NTT_{p_r}(limb plane of row d, component c)
int tggsw_precompute(Tggsw &g, const u64 *rows_dev, cudaStream_t st) {
    const TorusCtx &tc = *g.tc;
    const u32 n = (u32)tc.n, k1 = (u32)g.k + 1;
    const size_t polys = (size_t)k1 * 64 * k1;  // row polynomials
    FHE_CUDA_OK(cudaMalloc((void **)&g.R1, polys * 2 * n * sizeof(u64)));
    FHE_CUDA_OK(cudaMalloc((void **)&g.R2, polys * 2 * n * sizeof(u64)));
    for (int r = 0; r < 2; r++) {
        u64 *R = r == 0 ? g.R1 : g.R2;
        split32_mod_kernel<<<grid_for(polys * n), 256, 0, st>>>(rows_dev, R, polys, n, r == 0 ? tc.cp.p1 : tc.cp.p2);
        count_launch(1);
        int rc = tc.ntt(r, MODE_FWD, R, R, polys * 2, st);
        if (rc) return rc;
    }
    FHE_CUDA_OK(cudaStreamSynchronize(st));
    return tggsw_build_fused_layout(g, st);
}

// ---------------------------------------------------------------------------------------------------
// TGGSW::encrypt_s on the device (tggsw.rs:17-33,100-122, glwe.rs:140-156 with R = Tn), counter-based
// sampler of the CPU restatement (orc_tggsw_encrypt_s_ctr): draw p of row r is SplitMix64 output r*(k*n + 12n) + p + 1.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 tg_draw(u64 seed, u64 pos) {
    u64 z = seed + (pos + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double tg_unit(u64 v) { return __dmul_rn(__ull2double_rn(v >> 11), 1.0 / 9007199254740992.0); }
// A[(r*k + c)*n + x] = mask coefficient, S[(r*k + c)*n + x] = sk_c[x] (operands of the batched Tn product);
// also negs[c*n + x] = -sk_c[x] and mrep[c*n + x] = m[x] for the messages -s_c * m
__global__ void tggsw_gen_masks_kernel(u64 *__restrict__ A, u64 *__restrict__ S, u64 *__restrict__ negs, u64 *__restrict__ mrep,
                                       const u64 *__restrict__ sk, const u64 *__restrict__ m, u64 seed, u32 n, u32 k, u32 rows,
                                       int uniform_mask) {
    const size_t kn = (size_t)k * n, per_row = kn + 12 * (size_t)n, total = (size_t)rows * kn;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t r = idx / kn, p = idx % kn;
        const u64 v = tg_draw(seed, r * per_row + p);
        A[idx] = uniform_mask ? v : __double2ull_rz(round(__dmul_rn(2.0, tg_unit(v))));
        S[idx] = sk[p];
        if (r == 0) {
            negs[p] = (u64)0 - sk[p];
            mrep[p] = m[p % n];
        }
    }
}
// rows[r] = (A_r, sum_c P_{r,c} + mi_{r/64} * g_lv + e)
// msgs != nullptr: plain batch encryption, row r = TGLWE_sk(msgs[r]) (TGLWE::encrypt_s, tglwe.rs:76-79)
__global__ void tggsw_gen_finish_kernel(u64 *__restrict__ out, const u64 *__restrict__ A, const u64 *__restrict__ P,
                                        const u64 *__restrict__ mi, u64 seed, u32 n, u32 k, u32 rows, double sigma,
                                        const u64 *__restrict__ msgs) {
    const size_t kn = (size_t)k * n, per_row = kn + 12 * (size_t)n, glwe = kn + n, total = (size_t)rows * glwe;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t r = idx / glwe, q = idx % glwe;
        if (q < kn) { out[idx] = A[r * kn + q]; continue; }
        const u32 x = (u32)(q - kn), i = (u32)(r / 64), lv = (u32)(r % 64) + 1;
        u64 b = 0;
        for (u32 c = 0; c < k; c++) b += P[(r * k + c) * n + x];
        double acc = 0.0;
        for (u32 t = 0; t < 12; t++) acc = __dadd_rn(acc, tg_unit(tg_draw(seed, r * per_row + kn + 12 * (size_t)x + t)));
        const u64 g = lv < 64 ? ~0ull / (1ull << lv) : 1ull;
        const u64 msg = msgs ? msgs[r * n + x] : mi[(size_t)i * n + x] * g;
        out[idx] = b + msg + __double2ull_rz(round(__dmul_rn(sigma, __dadd_rn(acc, -6.0))));
    }
}
// rows_out: (k+1)*64 TGLWEs of (k+1)*n words, device pointer
int tggsw_generate_device(const TorusCtx &tc, u64 k, const u64 *sk, const u64 *m, double sigma, u64 seed, int uniform_mask,
                          u64 *rows_out, cudaStream_t st) {
    const u32 n = (u32)tc.n, rows = (u32)((k + 1) * 64);
    const size_t kn = (size_t)k * n, words = (size_t)rows * kn;
    Scratch sA, sS, sP, sN, sM, sMi;
    int rc;
    if ((rc = sA.alloc(words * 8, st)) || (rc = sS.alloc(words * 8, st)) || (rc = sP.alloc(words * 8, st)) ||
        (rc = sN.alloc(kn * 8, st)) || (rc = sM.alloc(kn * 8, st)) || (rc = sMi.alloc((kn + n) * 8, st)))
        return rc;
    u64 *A = sA.ptr<u64>(), *S = sS.ptr<u64>(), *P = sP.ptr<u64>(), *mi = sMi.ptr<u64>();
    tggsw_gen_masks_kernel<<<grid_for(words), 256, 0, st>>>(A, S, sN.ptr<u64>(), sM.ptr<u64>(), sk, m, seed, n, (u32)k, rows,
                                                          uniform_mask);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    if ((rc = tn_mul_device(tc, sN.ptr<u64>(), sM.ptr<u64>(), mi, k, st))) return rc;          // mi_c = -s_c * m  (tggsw.rs:28)
    FHE_CUDA_OK(cudaMemcpyAsync(mi + kn, m, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));       // mi_k = m
    if ((rc = tn_mul_device(tc, A, S, P, (size_t)rows * k, st))) return rc;                     // a_{r,c} * s_c
    tggsw_gen_finish_kernel<<<grid_for((size_t)rows * (kn + n)), 256, 0, st>>>(rows_out, A, P, mi, seed, n, (u32)k, rows, sigma,
                                                                              nullptr);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}
// out[b] = TGLWE_sk(msgs[b]) for `batch` message polynomials (same sampler: ciphertext b uses the draws of row b)
int tglwe_encrypt_device(const TorusCtx &tc, u64 k, const u64 *sk, const u64 *msgs, size_t batch, double sigma, u64 seed,
                         int uniform_mask, u64 *out, cudaStream_t st) {
    const u32 n = (u32)tc.n;
    FHE_REQUIRE(batch < (1ull << 31), "tglwe encrypt: batch too large");
    const size_t kn = (size_t)k * n, words = batch * kn;
    Scratch sA, sS, sP, sN, sM;
    int rc;
    if ((rc = sA.alloc(words * 8, st)) || (rc = sS.alloc(words * 8, st)) || (rc = sP.alloc(words * 8, st)) ||
        (rc = sN.alloc(kn * 8, st)) || (rc = sM.alloc(kn * 8, st)))
        return rc;
    tggsw_gen_masks_kernel<<<grid_for(words), 256, 0, st>>>(sA.ptr<u64>(), sS.ptr<u64>(), sN.ptr<u64>(), sM.ptr<u64>(), sk, msgs, seed,
                                                          n, (u32)k, (u32)batch, uniform_mask);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    if ((rc = tn_mul_device(tc, sA.ptr<u64>(), sS.ptr<u64>(), sP.ptr<u64>(), batch * k, st))) return rc;
    tggsw_gen_finish_kernel<<<grid_for(batch * (kn + n)), 256, 0, st>>>(out, sA.ptr<u64>(), sP.ptr<u64>(), nullptr, seed, n, (u32)k,
                                                                       (u32)batch, sigma, msgs);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

// TGLWE::decrypt (tglwe.rs:86-88 -> glwe.rs:175-179 with R = Tn): p_b = ct_b.b - sum_i ct_b.a_i * sk_i
__global__ void tglwe_dec_gather_kernel(const u64 *__restrict__ ct, const u64 *__restrict__ sk, u64 *__restrict__ A,
                                        u64 *__restrict__ S, size_t batch, u32 n, u32 k) {
    const size_t kn = (size_t)k * n, total = batch * kn;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx / kn, p = idx % kn;
        A[idx] = ct[b * (kn + n) + p];
        S[idx] = sk[p];
    }
}
__global__ void tglwe_dec_finish_kernel(const u64 *__restrict__ ct, const u64 *__restrict__ P, u64 *__restrict__ out,
                                        size_t batch, u32 n, u32 k) {
    const size_t kn = (size_t)k * n, total = batch * n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx / n;
        const u32 x = (u32)(idx % n);
        u64 acc = 0;
        for (u32 i = 0; i < k; i++) acc += P[b * kn + (size_t)i * n + x];
        out[idx] = ct[b * (kn + n) + kn + x] - acc;
    }
}
int tglwe_decrypt_device(const TorusCtx &tc, u64 k, const u64 *sk, const u64 *ct, u64 *out, size_t batch, cudaStream_t st) {
    const u32 n = (u32)tc.n;
    const size_t words = batch * k * n;
    Scratch sA, sS, sP;
    int rc;
    if ((rc = sA.alloc(words * 8, st)) || (rc = sS.alloc(words * 8, st)) || (rc = sP.alloc(words * 8, st))) return rc;
    tglwe_dec_gather_kernel<<<grid_for(words), 256, 0, st>>>(ct, sk, sA.ptr<u64>(), sS.ptr<u64>(), batch, n, (u32)k);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    if ((rc = tn_mul_device(tc, sA.ptr<u64>(), sS.ptr<u64>(), sP.ptr<u64>(), batch * k, st))) return rc;
    tglwe_dec_finish_kernel<<<grid_for(batch * n), 256, 0, st>>>(ct, sP.ptr<u64>(), out, batch, n, (u32)k);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

int tn_addsub_device(const u64 *a, const u64 *b, u64 *c, size_t len, int op, cudaStream_t st) {
    tn_addsub_kernel<<<grid_for(len), 256, 0, st>>>(a, b, c, len, op);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}
int cmux_chain_device(const Tggsw *const *gs, size_t steps, const u64 *acc_in, const u64 *h, int negacyclic, u64 *acc_out,
                      size_t batch, cudaStream_t st) {
    FHE_REQUIRE(steps <= 0x7fffffffull, "cmux chain: too many steps");
    FHE_REQUIRE(steps >= 1, "cmux chain: internal, zero steps are handled by the caller");
    const TorusCtx &tc = *gs[0]->tc;
    const u32 n = (u32)tc.n, k1 = (u32)gs[0]->k + 1;
    const size_t glwe = (size_t)k1 * n;
    bool all_fused = true;
    for (size_t j = 0; j < steps; j++) {
        FHE_REQUIRE(gs[j]->tc == &tc && gs[j]->k + 1 == k1, "cmux chain: every TGGSW must share (n, k)");
        all_fused = all_fused && gs[j]->R1f != nullptr;
    }
    const char *force = getenv("FHE_EXTPROD_PATH");
    if (force && strncmp(force, "fused", 5) == 0 && !all_fused) {
        set_error("FHE_EXTPROD_PATH=fused but this (n, k) has no fused instantiation");
        return -1;
    }
    int rc;
    if (all_fused && !(force && strcmp(force, "unfused") == 0)) {
        std::vector<const u32 *> keys(2 * steps);
        for (size_t j = 0; j < steps; j++) {
            keys[2 * j] = gs[j]->R1f;
            keys[2 * j + 1] = gs[j]->R2f;
        }
        Scratch s_keys;
        if ((rc = s_keys.alloc(keys.size() * sizeof(void *), st))) return rc;
        // pageable source: the copy is staged before the call returns, so `keys` may die with this frame
        FHE_CUDA_OK(cudaMemcpyAsync(s_keys.ptr<void>(), keys.data(), keys.size() * sizeof(void *), cudaMemcpyHostToDevice, st));
        return cmux_chain_fused_device(tc, (int)k1, s_keys.ptr<const u32 *>(), h, (int)steps, negacyclic, acc_in, acc_out,
                                       batch, st);
    }
    Scratch s_rot, s_a, s_b;
    if ((rc = s_rot.alloc(batch * glwe * 8, st))) return rc;
    if ((rc = s_a.alloc(batch * glwe * 8, st))) return rc;
    if ((rc = s_b.alloc(batch * glwe * 8, st))) return rc;
    u64 *rot = s_rot.ptr<u64>(), *pp[2] = {s_a.ptr<u64>(), s_b.ptr<u64>()};
    const u64 *cur = acc_in;
    for (size_t j = 0; j < steps; j++) {
        if ((rc = tn_left_rotate_device(cur, rot, batch * k1, n, h + j, 0, k1, st, steps, negacyclic))) return rc;
        if ((rc = extprod_device(*gs[j], cur, rot, pp[j & 1], batch, st))) return rc;  // cmux(gs[j], acc, rot)
        cur = pp[j & 1];
    }
    FHE_CUDA_OK(cudaMemcpyAsync(acc_out, cur, batch * glwe * 8, cudaMemcpyDeviceToDevice, st));
    return 0;
}
int tn_left_rotate_device(const u64 *a, u64 *out, size_t polys, u32 n, const u64 *hs, u64 h_const, u32 group,
                          cudaStream_t st, size_t hs_stride, int negacyclic) {
    tn_left_rotate_kernel<<<grid_for(polys * n), 256, 0, st>>>(a, out, polys, n, hs, h_const, group, hs_stride, negacyclic);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fhe
