// ks_mma.cu -- tensor-core key switch (the one GEMM-shaped operation on the path).
//
//   out[b][x] = lhs[b][x] - sum_{r < kn_in*64} A[b][r] * KSK[r][x]   (mod 2^64),  A[b][i*64+bp] = bit bp of a_{b,i}
// (tfhe/src/tlwe.rs:101-112 with beta=2, l=64: digit j of T64::decompose is bit 63-j, torus.rs:43-52).
// The u64 key is split into its 8 byte planes, KSK[r][x] = sum_p 2^(8p) B[r][x*8+p], so the product is an
// exact u8 x u8 -> s32 GEMM  S[b][c] = sum_r A[b][r] * B[r][c]  (S <= 65536*255 < 2^24 for kn_in <= 1024;
// checked at load), recombined as sum_p S[b][x*8+p] << 8p (mod 2^64).  M = batch, N = 8*(kn_out+1), K = kn_in*64.
//
// Kernel: CTA tile 128 (ciphertexts) x 256 (columns = 32 key words x 8 planes), K step 128.
//   * B: the key is re-laid out ONCE at load into 32 KB blocks [n_tile][k_tile] that are the exact
//     (XOR-swizzled, ldmatrix conflict-free) shared-memory image, so a stage is one cp.async.bulk (TMA 1-D)
//     with mbarrier completion; 4-stage ring.
//   * A: never materialised in HBM.  Each thread expands one 64-bit mask word per K step into 64 bytes
//     (nibble * 0x00204081 & 0x01010101) straight into the swizzled A tile (double buffered).
//   * 8 warps as 2 (M) x 4 (N), warp tile 64 x 64 = 4 x 8 mma.sync.m16n8k32 (u8.u8.s32) per 32-wide k
//     slice, fragments through ldmatrix.x4.  One __syncthreads per K step.
//   * Epilogue: an n-tile of 8 columns is exactly one key word; the 8 plane sums are shifted, reduced over
//     the 4 lanes of a quad by shuffles and subtracted from (0,..,0,b) (tlwe.rs:111).
#include <algorithm>

#include "../../include/fhe_b200.h"
#include "runtime.cuh"
#include "tlwe.cuh"

namespace fhe {

constexpr int KM_BM = 128, KM_BN = 256, KM_BK = 128, KM_STAGES = 4, KM_THREADS = 256;
constexpr int KM_BBYTES = KM_BN * KM_BK;  // 32 KB per B stage
constexpr int KM_ABYTES = KM_BM * KM_BK;  // 16 KB per A buffer
constexpr size_t KM_SMEM = (size_t)KM_STAGES * KM_BBYTES + 2 * KM_ABYTES + 64;

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u32 swz(u32 row, u32 chunk) { return row * 128u + ((chunk ^ (row & 7u)) << 4); }

__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void ldsm_x4(u32 addr, u32 &r0, u32 &r1, u32 &r2, u32 &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_u8(int (&c)[4], u32 a0, u32 a1, u32 a2, u32 a3, u32 b0, u32 b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 4 bits -> 4 bytes (byte i = bit i)
__device__ __forceinline__ u32 spread4(u32 nib) { return (nib * 0x00204081u) & 0x01010101u; }

// ---------------------------------------------------------------------------------------------------
// one-time key re-layout: rows[(i*64 + j)*w + x]  ->  blocks[(nt*KT + kt)][row = xl*8 + p][k = ih*64 + bp]
// (swizzled shared-memory image), bp = 63 - j, i = 2*kt + ih, x = nt*32 + xl; zero padding for x >= w.
// ---------------------------------------------------------------------------------------------------
__global__ void ksk_relayout_kernel(const u64 *__restrict__ rows, unsigned char *__restrict__ blocks, u32 kn_in, u32 w,
                                    u32 n_tiles) {
    const u32 KT = kn_in / 2;
    const size_t total = (size_t)n_tiles * KT * (KM_BBYTES / 16);  // one thread per 16-byte chunk
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const u32 chunk_phys = (u32)(idx % 8), row = (u32)((idx / 8) % KM_BN);
        const size_t blk = idx / (8 * KM_BN);
        const u32 kt = (u32)(blk % KT), nt = (u32)(blk / KT);
        const u32 chunk = chunk_phys ^ (row & 7u);
        const u32 xl = row >> 3, p = row & 7u, x = nt * 32 + xl;
        const u32 i = 2 * kt + (chunk >> 2);
        const u32 bp0 = (chunk & 3u) * 16;
        u32 v[4] = {0, 0, 0, 0};
        if (x < w) {
#pragma unroll
            for (u32 t = 0; t < 16; t++) {
                const u32 bp = bp0 + t, j = 63 - bp;
                const u64 word = rows[((size_t)i * 64 + j) * w + x];
                v[t >> 2] |= (u32)((word >> (8 * p)) & 0xffull) << (8 * (t & 3u));
            }
        }
        reinterpret_cast<uint4 *>(blocks)[idx] = make_uint4(v[0], v[1], v[2], v[3]);
    }
}

// ---------------------------------------------------------------------------------------------------
// the GEMM
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KM_THREADS, 1)
ks_mma_kernel(const unsigned char *__restrict__ blocks, const u64 *__restrict__ ct, u64 *__restrict__ out, size_t batch,
              u32 kn_in, u32 kn_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sB = smem;
    unsigned char *sA = smem + (size_t)KM_STAGES * KM_BBYTES;
    u64 *bars = reinterpret_cast<u64 *>(sA + 2 * KM_ABYTES);
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps
    const u32 KT = kn_in / 2;
    const u32 w = kn_out + 1;
    const size_t b0 = (size_t)blockIdx.x * KM_BM;  // m-tile fastest: CTAs sharing a key slab are co-resident
    const u32 nt = blockIdx.y;
    const unsigned char *gB = blocks + (size_t)nt * KT * KM_BBYTES;

    if (tid == 0) {
        for (int s = 0; s < KM_STAGES; s++) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (u32 t = 0; t + 1 < KM_STAGES && t < KT; t++) {
            mbar_expect_tx(smem_u32(&bars[t]), KM_BBYTES);
            bulk_g2s(smem_u32(sB + (size_t)t * KM_BBYTES), gB + (size_t)t * KM_BBYTES, KM_BBYTES, smem_u32(&bars[t]));
        }
    }

    // A producer role of this thread: row tid/2, mask word (2*kt + tid%2)
    const u32 arow = tid >> 1, ah = tid & 1;
    const bool arow_ok = (b0 + arow) < batch;
    const u64 *actp = ct + (b0 + arow) * (size_t)(kn_in + 1) + ah;
    u64 w_next = arow_ok ? __ldg(actp) : 0;

    int acc[4][8][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++)
#pragma unroll
            for (int e = 0; e < 4; e++) acc[i][j][e] = 0;

    // ldmatrix lane roles
    const u32 mi = lane >> 3, r8 = lane & 7;
    const u32 a_row = wm * 64 + (mi & 1) * 8 + r8;   // + mt*16
    const u32 a_chk = mi >> 1;                       // + 2*ks
    const u32 b_row = wn * 64 + (mi >> 1) * 8 + r8;  // + np*16
    const u32 b_chk = mi & 1;                        // + 2*ks

    for (u32 kt = 0; kt < KT; kt++) {
        const u32 stage = kt % KM_STAGES, parity = (kt / KM_STAGES) & 1;
        const u64 wd = w_next;
        if (kt + 1 < KT) w_next = arow_ok ? __ldg(actp + 2 * (size_t)(kt + 1)) : 0;
        unsigned char *A = sA + (size_t)(kt & 1) * KM_ABYTES;
        {   // expand 64 bits -> 64 bytes -> 4 swizzled 16-byte chunks of row `arow`
            const u32 lo = (u32)wd, hi = (u32)(wd >> 32);
#pragma unroll
            for (u32 c = 0; c < 4; c++) {
                const u32 half = (c < 2 ? lo : hi) >> ((c & 1) * 16);
                uint4 v;
                v.x = spread4(half & 15u);
                v.y = spread4((half >> 4) & 15u);
                v.z = spread4((half >> 8) & 15u);
                v.w = spread4((half >> 12) & 15u);
                *reinterpret_cast<uint4 *>(A + swz(arow, ah * 4 + c)) = v;
            }
        }
        mbar_wait(smem_u32(&bars[stage]), parity);
        __syncthreads();
        if (tid == 0 && kt + KM_STAGES - 1 < KT) {  // refill the stage that iteration kt-1 finished reading
            const u32 t = kt + KM_STAGES - 1, s = t % KM_STAGES;
            mbar_expect_tx(smem_u32(&bars[s]), KM_BBYTES);
            bulk_g2s(smem_u32(sB + (size_t)s * KM_BBYTES), gB + (size_t)t * KM_BBYTES, KM_BBYTES, smem_u32(&bars[s]));
        }
        const u32 sAaddr = smem_u32(A), sBaddr = smem_u32(sB + (size_t)stage * KM_BBYTES);
#pragma unroll
        for (u32 ks = 0; ks < KM_BK / 32; ks++) {
            u32 af[4][4];
#pragma unroll
            for (u32 mt = 0; mt < 4; mt++)
                ldsm_x4(sAaddr + swz(a_row + mt * 16, 2 * ks + a_chk), af[mt][0], af[mt][1], af[mt][2], af[mt][3]);
#pragma unroll
            for (u32 np = 0; np < 4; np++) {
                u32 bf[4];
                ldsm_x4(sBaddr + swz(b_row + np * 16, 2 * ks + b_chk), bf[0], bf[1], bf[2], bf[3]);
#pragma unroll
                for (u32 mt = 0; mt < 4; mt++) {
                    mma_u8(acc[mt][2 * np], af[mt][0], af[mt][1], af[mt][2], af[mt][3], bf[0], bf[1]);
                    mma_u8(acc[mt][2 * np + 1], af[mt][0], af[mt][1], af[mt][2], af[mt][3], bf[2], bf[3]);
                }
            }
        }
    }

    // epilogue: n-tile j of this warp is key word x = nt*32 + wn*8 + j; thread holds planes 2*tig, 2*tig+1
    const u32 g = lane >> 2, tig = lane & 3;
#pragma unroll
    for (u32 mt = 0; mt < 4; mt++) {
#pragma unroll
        for (u32 j = 0; j < 8; j++) {
#pragma unroll
            for (u32 h = 0; h < 2; h++) {  // rows g and g+8
                u64 s = ((u64)(u32)acc[mt][j][2 * h] << (16 * tig)) + ((u64)(u32)acc[mt][j][2 * h + 1] << (16 * tig + 8));
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                const size_t b = b0 + wm * 64 + mt * 16 + h * 8 + g;
                const u32 x = nt * 32 + wn * 8 + j;
                if (tig == 0 && b < batch && x < w) {
                    const u64 lhs = x == kn_out ? ct[b * (size_t)(kn_in + 1) + kn_in] : 0;
                    out[b * (size_t)w + x] = lhs - s;
                }
            }
        }
    }
}

__global__ void ksk_gather_bcol_kernel(const u64 *__restrict__ rows, u64 *__restrict__ bcol, size_t nrows, u32 w) {
    for (size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x; r < nrows; r += (size_t)gridDim.x * blockDim.x)
        bcol[r] = rows[r * w + (w - 1)];
}

int ksk_build_mma_layout(Ksk &k, cudaStream_t st) {
    const u32 w = (u32)k.kn_out + 1;
    const bool split_b = k.kn_out % 32 == 0;  // see tlwe.cuh: Ksk::bcol
    k.mma_n_tiles = split_b ? (u32)(k.kn_out / 32) : (w + 31) / 32;
    if (split_b) {
        const size_t nrows = (size_t)k.kn_in * k.l;
        FHE_CUDA_OK(cudaMalloc((void **)&k.bcol, nrows * sizeof(u64)));
        ksk_gather_bcol_kernel<<<(unsigned)std::min<size_t>((nrows + 255) / 256, (size_t)num_sms() * 8), 256, 0, st>>>(k.rows, k.bcol, nrows, w);
        count_launch(1);
    }
    const size_t bytes = (size_t)k.mma_n_tiles * (k.kn_in / 2) * KM_BBYTES;
    FHE_CUDA_OK(cudaMalloc((void **)&k.mma_blocks, bytes));
    const size_t chunks = bytes / 16;
    size_t grid = (chunks + 255) / 256;
    if (grid > (size_t)num_sms() * 32) grid = (size_t)num_sms() * 32;
    ksk_relayout_kernel<<<(unsigned)grid, 256, 0, st>>>(k.rows, k.mma_blocks, (u32)k.kn_in, w, k.mma_n_tiles);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    FHE_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int key_switch_mma_device(const Ksk &k, const u64 *ct, u64 *out, size_t batch, cudaStream_t st) {
    static unsigned long long done_mask = 0;
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    if (!((done_mask >> (dev & 63)) & 1ull)) {
        FHE_CUDA_OK(cudaFuncSetAttribute(ks_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KM_SMEM));
        done_mask |= 1ull << (dev & 63);
    }
    dim3 grid((unsigned)((batch + KM_BM - 1) / KM_BM), k.mma_n_tiles);
    ks_mma_kernel<<<grid, KM_THREADS, KM_SMEM, st>>>(k.mma_blocks, ct, out, batch, (u32)k.kn_in, (u32)k.kn_out);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fhe
