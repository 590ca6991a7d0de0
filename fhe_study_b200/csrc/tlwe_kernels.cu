// tlwe_kernels.cu -- TLWE key switch and the bootstrapping pipeline as the reference EXECUTES it
// (tfhe/src/tlwe.rs:101-161): mod_switch -> one public left_rotate of the table TGLWE ->
// sample_extraction(0) -> key_switch(beta=2, l).  The key switch is a {0,1} x u64 matrix product mod 2^64:
//   out = (0,..,0, b) - sum_{i<kn_in} sum_{j<l} bit_{l-1-j}(a_i) * KSK[i][j][:]      (tlwe.rs:101-112,
//   tlev.rs:95-105, tlwe.rs:269-279, torus.rs:43-52)
// with a kn_in*l*(kn_out+1)*8-byte key resident in HBM (537 MB for n=1024, k=1, l=64).
#include <algorithm>

#include "../../include/fhe_b200.h"
#include "runtime.cuh"
#include "tlwe.cuh"

#include <stdlib.h>

namespace fhe {

// ---------------------------------------------------------------------------------------------------
// CUDA-core key switch.  CTA = KS_TX threads = KS_TX output columns; each thread keeps KS_BT accumulators
// (one per ciphertext of the CTA's batch tile) in registers and streams the KSK column slab once.
// ---------------------------------------------------------------------------------------------------
constexpr int KS_TX = 128;
constexpr int KS_BT = 16;
constexpr int KS_IC = 32;  // mask words staged per shared-memory refill

__global__ void __launch_bounds__(KS_TX)
key_switch_kernel(const u64 *__restrict__ ksk, const u64 *__restrict__ ct, u64 *__restrict__ out, size_t batch,
                  u32 kn_in, u32 kn_out, u32 l) {
    __shared__ u64 s_a[KS_IC][KS_BT];
    const u32 w = kn_out + 1;
    const u32 x = blockIdx.x * KS_TX + threadIdx.x;
    const size_t b0 = (size_t)blockIdx.y * KS_BT;
    const bool col_ok = x < w;
    u64 acc[KS_BT];
#pragma unroll
    for (int b = 0; b < KS_BT; b++) acc[b] = 0;
    for (u32 i0 = 0; i0 < kn_in; i0 += KS_IC) {
        __syncthreads();
        for (u32 t = threadIdx.x; t < KS_IC * KS_BT; t += KS_TX) {
            const u32 ii = t / KS_BT, b = t % KS_BT;
            const bool ok = (i0 + ii) < kn_in && (b0 + b) < batch;
            s_a[ii][b] = ok ? ct[(b0 + b) * (size_t)(kn_in + 1) + i0 + ii] : 0;
        }
        __syncthreads();
        const u32 ni = min((u32)KS_IC, kn_in - i0);
        for (u32 ii = 0; ii < ni; ii++) {
            u64 a[KS_BT];
#pragma unroll
            for (int b = 0; b < KS_BT; b++) a[b] = s_a[ii][b] << (64 - l);  // digit j is now bit (63-j)
            const u64 *row = ksk + ((size_t)(i0 + ii) * l) * w + x;
            for (u32 j = 0; j < l; j++) {
                const u64 v = col_ok ? __ldg(row + (size_t)j * w) : 0;
#pragma unroll
                for (int b = 0; b < KS_BT; b++) {
                    if ((i64)a[b] < 0) acc[b] += v;  // TLWE * T64(bit) summed (tlev.rs:95-105)
                    a[b] <<= 1;
                }
            }
        }
    }
    if (col_ok) {
#pragma unroll
        for (int b = 0; b < KS_BT; b++) {
            if (b0 + b < batch) {
                const u64 lhs = x == kn_out ? ct[(b0 + b) * (size_t)(kn_in + 1) + kn_in] : 0;  // (0,..,0,b)
                out[(b0 + b) * (size_t)w + x] = lhs - acc[b];                                    // tlwe.rs:111
            }
        }
    }
}

// Body column of the key switch when the GEMM covers the mask columns only (Ksk::bcol):
//   out[b][kn_out] = ct[b][kn_in] - sum_{i,j} bit_{l-1-j}(ct[b][i]) * bcol[i*l + j]      (tlwe.rs:101-112, column kn_out)
// Block = 32 ciphertexts (one per lane) x 8 warps striding the mask words of this block's K slice; the column
// entries are warp-uniform loads.  Partial sums are combined with 64-bit atomic adds (wrapping adds commute).
__global__ void __launch_bounds__(256)
ks_bcol_kernel(const u64 *__restrict__ bcol, const u64 *__restrict__ ct, u64 *__restrict__ out, size_t batch, u32 kn_in,
               u32 kn_out, u32 l) {
    __shared__ u64 red[8][32];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t b = (size_t)blockIdx.x * 32 + lane;
    const u32 per = (kn_in + gridDim.y - 1) / gridDim.y;
    const u32 i0 = blockIdx.y * per, i1 = min(kn_in, i0 + per);
    const u64 *cp = ct + b * (size_t)(kn_in + 1);
    u64 acc = 0;
    for (u32 i = i0 + warp; i < i1; i += 8) {
        const u64 a = b < batch ? __ldg(cp + i) : 0;
        const u64 *kb = bcol + (size_t)i * l;
#pragma unroll 8
        for (u32 j = 0; j < l; j++) acc += ((a >> (l - 1 - j)) & 1ull) ? __ldg(kb + j) : 0ull;
    }
    red[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && b < batch) {
        u64 s = 0;
#pragma unroll
        for (int w8 = 0; w8 < 8; w8++) s += red[w8][lane];
        const u64 v = (blockIdx.y == 0 ? __ldg(cp + kn_in) : 0ull) - s;
        atomicAdd(reinterpret_cast<unsigned long long *>(out + b * (size_t)(kn_out + 1) + kn_out), (unsigned long long)v);
    }
}
// Device-side KSK generation: structure of tlwe.rs:84-100 / tlev.rs:53-77 / glwe.rs:140-156
// (row i*l + lv-1 = TLWE_{new_sk}(sk_i * g_lv)), counter-based SplitMix64 sampler -- the one the CPU restatement
// (orc_tlwe_new_ksk_ctr) documents; every draw is addressed by (row, position), so rows are independent.
// One warp per row; f64 steps with explicit IEEE intrinsics (bit-identical to the CPU restatement).
__device__ __forceinline__ u64 ctr_draw(u64 seed, u64 pos) {
    u64 z = seed + (pos + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double ctr_unit(u64 v) { return __dmul_rn(__ull2double_rn(v >> 11), 1.0 / 9007199254740992.0); }
__device__ __forceinline__ u64 f64_as_u64_sat(double x) { return __double2ull_rz(x); }  // Rust `as u64`: saturating, NaN -> 0
__global__ void __launch_bounds__(256)
ksk_generate_kernel(u64 *__restrict__ rows, const u64 *__restrict__ sk, const u64 *__restrict__ new_sk, u64 seed, u32 kn_in,
                    u32 kn_out, u32 l, double sigma, int uniform_mask, const u64 *__restrict__ msgs, size_t nmsgs) {
    // msgs != nullptr: plain batch encryption, row r = TLWE_{new_sk}(msgs[r]) (TLWE::encrypt_s, tlwe.rs:71-74)
    const u32 lane = threadIdx.x & 31;
    const size_t nrows = msgs ? nmsgs : (size_t)kn_in * l, per_row = (size_t)kn_out + 12, w = (size_t)kn_out + 1;
    for (size_t r = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < nrows; r += (size_t)gridDim.x * 8) {
        const size_t base = r * per_row;
        u64 *row = rows + r * w;
        u64 part = 0;
        for (u32 x = lane; x < kn_out; x += 32) {
            const u64 v = ctr_draw(seed, base + x);
            const u64 a = uniform_mask ? v : f64_as_u64_sat(round(__dmul_rn(2.0, ctr_unit(v))));
            row[x] = a;
            part += a * new_sk[x];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) {
            double acc = 0.0;
            for (u32 t = 0; t < 12; t++) acc = __dadd_rn(acc, ctr_unit(ctr_draw(seed, base + kn_out + t)));
            u64 msg;
            if (msgs) {
                msg = msgs[r];
            } else {
                const u32 i = (u32)(r / l), lv = (u32)(r % l) + 1;
                msg = sk[i] * (lv < 64 ? ~0ull / (1ull << lv) : 1ull);
            }
            row[kn_out] = part + msg + f64_as_u64_sat(round(__dmul_rn(sigma, __dadd_rn(acc, -6.0))));
        }
    }
}
int ksk_generate_device(u64 *rows, const u64 *sk, const u64 *new_sk, u64 seed, u32 kn_in, u32 kn_out, u32 l, double sigma,
                        int uniform_mask, cudaStream_t st) {
    const size_t nrows = (size_t)kn_in * l;
    const unsigned grid = (unsigned)std::min<size_t>((nrows + 7) / 8, (size_t)num_sms() * 16);
    ksk_generate_kernel<<<grid, 256, 0, st>>>(rows, sk, new_sk, seed, kn_in, kn_out, l, sigma, uniform_mask, nullptr, 0);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}
int tlwe_encrypt_device(u64 *out, const u64 *sk, const u64 *msgs, size_t batch, u64 seed, u32 kn, double sigma, int uniform_mask,
                        cudaStream_t st) {
    const unsigned grid = (unsigned)std::min<size_t>((batch + 7) / 8, (size_t)num_sms() * 16);
    ksk_generate_kernel<<<grid, 256, 0, st>>>(out, nullptr, sk, seed, 0, kn, 1, sigma, uniform_mask, msgs, batch);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

int key_switch_bcol_device(const Ksk &k, const u64 *ct, u64 *out, size_t batch, cudaStream_t st) {
    const size_t w = k.kn_out + 1;
    FHE_CUDA_OK(cudaMemset2DAsync(out + k.kn_out, w * sizeof(u64), 0, sizeof(u64), batch, st));
    const unsigned ksplit = (unsigned)std::max<size_t>(1, std::min<size_t>(k.kn_in / 8, ((size_t)num_sms() * 8 * 32 + batch - 1) / batch));
    dim3 grid((unsigned)((batch + 31) / 32), std::min(ksplit, 64u));
    ks_bcol_kernel<<<grid, 256, 0, st>>>(k.bcol, ct, out, batch, (u32)k.kn_in, (u32)k.kn_out, (u32)k.l);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

int key_switch_device(const Ksk &k, const u64 *ct, u64 *out, size_t batch, cudaStream_t st) {
    // path selection (l == 64 keys carry the byte-plane tensor-core layout): tcgen05 GEMM for batches that fill
    // an MMA tile, the mma.sync GEMM for small batches, CUDA cores otherwise.  FHE_KS_PATH=tc|mma|cuda forces
    // one (the tests cover all three).
    const char *force = getenv("FHE_KS_PATH");
    int path = k.mma_blocks == nullptr ? 0 : batch >= 64 ? 2 : batch >= 16 ? 1 : 0;
    if (force) {
        path = strcmp(force, "tc") == 0 ? 2 : strcmp(force, "mma") == 0 ? 1 : 0;
        if (path != 0 && k.mma_blocks == nullptr) {
            set_error("FHE_KS_PATH asks for a tensor-core path but this key has no byte-plane layout (needs l == 64, even kn_in)");
            return -1;
        }
    }
    if (path != 0 && k.bcol != nullptr) {  // the GEMMs cover the mask columns only
        int rc = key_switch_bcol_device(k, ct, out, batch, st);
        if (rc) return rc;
    }
    if (path == 2) return key_switch_tc_device(k, ct, out, batch, st);
    if (path == 1) return key_switch_mma_device(k, ct, out, batch, st);
    const u32 w = (u32)k.kn_out + 1;
    dim3 grid((w + KS_TX - 1) / KS_TX, (unsigned)((batch + KS_BT - 1) / KS_BT));
    FHE_REQUIRE(grid.y <= 65535, "key switch: batch too large for one launch (max 65535*16)");
    key_switch_kernel<<<grid, KS_TX, 0, st>>>(k.rows, ct, out, batch, (u32)k.kn_in, (u32)k.kn_out, (u32)k.l);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// blind_rotation as executed + sample_extraction(0), fused (tlwe.rs:121-148, tglwe.rs:89-119):
//   h   = mod_switch(c.b, k*n) = c.b >> (64 - log2(k*n))                      (torus.rs:58-66)
//   acc = table.left_rotate(h)   : acc_i[c] = c+h' < n ? t_i[c+h'] : -t_i[c+h'-n], h' = h mod n
//   ext[n*i + j] = j <= 0 ? acc_i[0-j] : -acc_i[n-j] ;  ext[k*n] = acc_k[0]
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 rotated_coeff(const u64 *__restrict__ poly, u32 n, u32 h, u32 c) {
    const u32 s = c + h;
    return s < n ? poly[s] : (u64)0 - poly[s - n];
}
__global__ void rotate_extract_kernel(const u64 *__restrict__ table, const u64 *__restrict__ ct, u64 *__restrict__ ext,
                                      u64 *__restrict__ acc_out, size_t batch, u32 n, u32 k, u32 c_kn, u32 shift) {
    const u32 kn = k * n;
    const size_t total = batch * (size_t)(kn + 1);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx / (kn + 1);
        const u32 e = (u32)(idx % (kn + 1));
        const u64 body = ct[b * (size_t)(c_kn + 1) + c_kn];
        const u32 h = (u32)((shift >= 64 ? body : body >> shift) % n);
        u64 v;
        if (e == kn) {
            v = rotated_coeff(table + (size_t)k * n, n, h, 0);
        } else {
            const u32 i = e / n, j = e % n;
            v = j == 0 ? rotated_coeff(table + (size_t)i * n, n, h, 0)
                       : (u64)0 - rotated_coeff(table + (size_t)i * n, n, h, n - j);
        }
        ext[idx] = v;
    }
    if (acc_out != nullptr) {  // optional: the rotated accumulator itself (blind_rotation's return value)
        const size_t tot2 = batch * (size_t)(k + 1) * n;
        for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < tot2;
             idx += (size_t)gridDim.x * blockDim.x) {
            const size_t b = idx / ((size_t)(k + 1) * n);
            const u32 r = (u32)(idx % ((size_t)(k + 1) * n)), i = r / n, c = r % n;
            const u64 body = ct[b * (size_t)(c_kn + 1) + c_kn];
            const u32 h = (u32)((shift >= 64 ? body : body >> shift) % n);
            acc_out[idx] = rotated_coeff(table + (size_t)i * n, n, h, c);
        }
    }
}

// Start of a CMux-chain blind rotation (extension of tlwe.rs:121-148, see fhe_bootstrap_chain): per ciphertext
// acc0 = X^{-b'} table and the per-step rotation amounts hs[b][j] from the mod-switched mask.
//   mode 0 (as written): c' = c >> shift (mod_switch to k*n), left_rotate semantics (h mod n), hs = c'.a[j]
//   mode 1 (working PBS): mod_switch to 2n, true negacyclic rotation, hs = (2n - c'.a[j]) mod 2n
__global__ void chain_prepare_kernel(const u64 *__restrict__ table, const u64 *__restrict__ ct, u64 *__restrict__ acc0,
                                     u64 *__restrict__ hs, size_t batch, u32 n, u32 k, u32 c_kn, u32 steps, u32 shift,
                                     int mode) {
    const size_t glwe = (size_t)(k + 1) * n;
    const size_t tot = batch * glwe;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < tot; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx / glwe;
        const u32 r = (u32)(idx % glwe), i = r / n, c = r % n;
        const u64 body = ct[b * (size_t)(c_kn + 1) + c_kn];
        const u64 hb = shift >= 64 ? body : body >> shift;
        const u64 v = rotated_coeff(table + (size_t)i * n, n, (u32)(hb % n), c);
        acc0[idx] = (mode && ((hb / n) & 1)) ? (u64)0 - v : v;
    }
    const size_t toth = batch * (size_t)steps;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < toth; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx / steps;
        const u32 j = (u32)(idx % steps);
        const u64 a = ct[b * (size_t)(c_kn + 1) + j];
        const u64 am = shift >= 64 ? a : a >> shift;
        hs[idx] = mode ? (2ull * n - am) % (2ull * n) : am;
    }
}

// TGLWE::sample_extraction(h) (tglwe.rs:89-115) for `batch` TGLWEs, one common h
__global__ void sample_extract_kernel(const u64 *__restrict__ ct, u64 *__restrict__ out, size_t batch, u32 n, u32 k,
                                      u32 h) {
    const u32 kn = k * n;
    const size_t total = batch * (size_t)(kn + 1);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx / (kn + 1);
        const u32 e = (u32)(idx % (kn + 1));
        const u64 *g = ct + b * (size_t)(k + 1) * n;
        u64 v;
        if (e == kn) v = g[(size_t)k * n + h];
        else {
            const u32 i = e / n, j = e % n;
            v = j <= h ? g[(size_t)i * n + (h - j)] : (u64)0 - g[(size_t)i * n + (n + h - j)];
        }
        out[idx] = v;
    }
}
// TLWE::mod_switch (tlwe.rs:114-118): every element >> shift
__global__ void shift_right_kernel(const u64 *__restrict__ a, u64 *__restrict__ out, size_t len, u32 shift) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x)
        out[i] = shift >= 64 ? a[i] : a[i] >> shift;
}

static inline unsigned grid_for(size_t work, int threads = 256) {
    size_t g = (work + threads - 1) / threads;
    const size_t cap = (size_t)num_sms() * 16;
    return (unsigned)(g < 1 ? 1 : g > cap ? cap : g);
}

int rotate_extract_device(const u64 *table, const u64 *ct, u64 *ext, u64 *acc_out, size_t batch, u32 n, u32 k, u32 c_kn,
                          cudaStream_t st) {
    const u32 log2kn = 63 - __builtin_clzll((u64)k * n);
    const u32 shift = 64 - log2kn;  // torus.rs:58-66 (release-mode shift; kn >= 2 in every caller)
    rotate_extract_kernel<<<grid_for(batch * (size_t)((k + 1) * n)), 256, 0, st>>>(table, ct, ext, acc_out, batch, n, k,
                                                                                  c_kn, shift);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}
// TLWE::decrypt (tlwe.rs:80-82 -> glwe.rs:175-179 with R = T64): phase_b = ct_b.b - <ct_b.a, sk>, one warp per ciphertext
__global__ void __launch_bounds__(256)
tlwe_decrypt_kernel(const u64 *__restrict__ sk, const u64 *__restrict__ ct, u64 *__restrict__ out, size_t batch, u32 kn) {
    const u32 lane = threadIdx.x & 31;
    for (size_t b = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5); b < batch; b += (size_t)gridDim.x * 8) {
        const u64 *c = ct + b * (size_t)(kn + 1);
        u64 part = 0;
        for (u32 i = lane; i < kn; i += 32) part += c[i] * sk[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) out[b] = c[kn] - part;
    }
}
int tlwe_decrypt_device(const u64 *sk, const u64 *ct, u64 *out, size_t batch, u32 kn, cudaStream_t st) {
    const unsigned grid = (unsigned)std::min<size_t>((batch + 7) / 8, (size_t)num_sms() * 16);
    tlwe_decrypt_kernel<<<grid, 256, 0, st>>>(sk, ct, out, batch, kn);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

int chain_prepare_device(const u64 *table, const u64 *ct, u64 *acc0, u64 *hs, size_t batch, u32 n, u32 k, u32 c_kn,
                         u32 steps, int mode, cudaStream_t st) {
    const u64 q2 = mode ? 2ull * n : (u64)k * n;
    const u32 shift = 64 - (63 - __builtin_clzll(q2));  // torus.rs:58-66
    chain_prepare_kernel<<<grid_for(batch * (size_t)((k + 1) * n)), 256, 0, st>>>(table, ct, acc0, hs, batch, n, k, c_kn,
                                                                                 steps, shift, mode);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}
int sample_extract_device(const u64 *ct, u64 *out, size_t batch, u32 n, u32 k, u32 h, cudaStream_t st) {
    sample_extract_kernel<<<grid_for(batch * (size_t)(k * n + 1)), 256, 0, st>>>(ct, out, batch, n, k, h);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}
int shift_right_device(const u64 *a, u64 *out, size_t len, u32 shift, cudaStream_t st) {
    shift_right_kernel<<<grid_for(len), 256, 0, st>>>(a, out, len, shift);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fhe
