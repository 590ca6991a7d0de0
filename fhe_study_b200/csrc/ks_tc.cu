// ks_tc.cu -- key switch on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
// Same exact byte-plane GEMM as ks_mma.cu (S[b][c] = sum_r A[b][r] * B[r][c], A = bits of the TLWE mask,
// B = u8 planes of the key, s32 accumulation, recombined mod 2^64), same pre-swizzled 32 KB key blocks
// (they are K-major SWIZZLE_128B tiles: 256 rows x 128 B, 8-row groups 1024 B apart).  What changes is the
// engine: the legacy mma.sync path saturates its IMMA pipe at ~0.8 POPS (profiles/r1_keyswitch_imma_*),
// tcgen05 reads both operands from shared memory through descriptors and accumulates in tensor memory.
//
// CTA = 256 ciphertexts x 256 columns (32 key words x 8 planes), K step 128, 3-stage ring (64 KB/stage):
//   warp 0      : one lane streams the key blocks with cp.async.bulk (TMA 1-D) -> full_b[s]
//   warp 1      : allocates TMEM (512 columns = two 128x256 s32 accumulators); one lane issues, per K step,
//                 4 x 2 tcgen05.mma (M=128, N=256, K=32) and commits them to empty[s]
//   warps 2..9  : 256 threads, one ciphertext row each: expand 2 mask words per K step into 128 bytes of the
//                 swizzled A tiles (never materialised in HBM), fence.proxy.async, arrive on full_a[s];
//                 at the end they are the epilogue: tcgen05.ld the accumulator rows, shift-and-add the 8 plane
//                 sums of each key word, subtract from (0,..,0,b) (tfhe/src/tlwe.rs:111) and store.
#include "../../include/fhe_b200.h"
#include "runtime.cuh"
#include "tlwe.cuh"

namespace fhe {

constexpr int TC_BM = 256, TC_BN = 256, TC_BK = 128, TC_STAGES = 3;
constexpr int TC_THREADS = 320;                       // 10 warps
constexpr int TC_A_HALF = 128 * TC_BK;                // 16 KB: one 128-row A tile
constexpr int TC_B_BYTES = TC_BN * TC_BK;             // 32 KB
constexpr int TC_STAGE = 2 * TC_A_HALF + TC_B_BYTES;  // 64 KB
constexpr size_t TC_SMEM = (size_t)TC_STAGES * TC_STAGE + 256;

__device__ __forceinline__ u32 tc_smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(u32 bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TC_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TC_DONE;\n"
        "bra TC_WAIT;\n"
        "TC_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tc_bulk_g2s(u32 dst, const void *src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, rows 128 B apart, 8-row groups 1024 B apart
__device__ __forceinline__ u64 tc_smem_desc(u32 saddr) {
    return (u64)((saddr >> 4) & 0x3FFFu) | ((u64)1 << 16) /* LBO (unused for swizzled K-major) */ |
           ((u64)(1024 >> 4) << 32) /* SBO */ | ((u64)1 << 46) /* descriptor version (sm_100) */ |
           ((u64)2 << 61) /* SWIZZLE_128B */;
}
// instruction descriptor: D = s32, A = B = u8, both K-major, M = 128, N = 256
constexpr u32 TC_IDESC = (2u << 4) | (0u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void tc_mma_i8(u32 tmem_d, u64 adesc, u64 bdesc, u32 accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(u32 bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ u32 tc_spread4(u32 nib) { return (nib * 0x00204081u) & 0x01010101u; }

__global__ void __launch_bounds__(TC_THREADS, 1)
ks_tc_kernel(const unsigned char *__restrict__ blocks, const u64 *__restrict__ ct, u64 *__restrict__ out, size_t batch,
             u32 kn_in, u32 kn_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    u64 *bars = reinterpret_cast<u64 *>(smem + (size_t)TC_STAGES * TC_STAGE);
    // bars[0..S) full_a, [S..2S) full_b, [2S..3S) empty, [3S] accumulators ready ; then the TMEM base address
    u32 *tmem_slot = reinterpret_cast<u32 *>(bars + 3 * TC_STAGES + 1);
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u32 KT = kn_in / 2, w = kn_out + 1;
    const size_t b0 = (size_t)blockIdx.x * TC_BM;
    const u32 nt = blockIdx.y;
    const unsigned char *gB = blocks + (size_t)nt * KT * TC_B_BYTES;
    auto full_a = [&](u32 s) { return tc_smem_u32(&bars[s]); };
    auto full_b = [&](u32 s) { return tc_smem_u32(&bars[TC_STAGES + s]); };
    auto empty = [&](u32 s) { return tc_smem_u32(&bars[2 * TC_STAGES + s]); };
    const u32 acc_bar = tc_smem_u32(&bars[3 * TC_STAGES]);

    if (tid == 0) {
        for (u32 s = 0; s < TC_STAGES; s++) {
            tc_mbar_init(full_a(s), 256);
            tc_mbar_init(full_b(s), 1);
            tc_mbar_init(empty(s), 1);
        }
        tc_mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: all 512 columns (two 128 x 256 s32 accumulators)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== key-block producer =====
        if (lane == 0) {
            for (u32 kt = 0; kt < KT; kt++) {
                const u32 s = kt % TC_STAGES, ph = (kt / TC_STAGES) & 1;
                tc_mbar_wait(empty(s), ph ^ 1);
                tc_mbar_expect_tx(full_b(s), TC_B_BYTES);
                tc_bulk_g2s(tc_smem_u32(smem + (size_t)s * TC_STAGE + 2 * TC_A_HALF), gB + (size_t)kt * TC_B_BYTES, TC_B_BYTES,
                            full_b(s));
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            for (u32 kt = 0; kt < KT; kt++) {
                const u32 s = kt % TC_STAGES, ph = (kt / TC_STAGES) & 1;
                tc_mbar_wait(full_a(s), ph);
                tc_mbar_wait(full_b(s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const u32 sa = tc_smem_u32(smem + (size_t)s * TC_STAGE);
#pragma unroll
                for (u32 k = 0; k < TC_BK / 32; k++) {
                    const u64 bd = tc_smem_desc(sa + 2 * TC_A_HALF + k * 32);
                    const u32 accum = (kt | k) != 0 ? 1u : 0u;
                    tc_mma_i8(tmem_base, tc_smem_desc(sa + k * 32), bd, accum);
                    tc_mma_i8(tmem_base + 256, tc_smem_desc(sa + TC_A_HALF + k * 32), bd, accum);
                }
                tc_commit(empty(s));  // arrives when the MMAs above have finished reading stage s
            }
            tc_commit(acc_bar);
        }
    } else {
        // ===== A expanders (one ciphertext row per thread), then the epilogue =====
        const u32 e = tid - 64;  // 0..255
        const u32 half = e >> 7, row = e & 127;
        const bool row_ok = (b0 + e) < batch;
        const u64 *cp = ct + (b0 + e) * (size_t)(kn_in + 1);
        u64 w0 = row_ok ? __ldg(cp) : 0, w1 = row_ok ? __ldg(cp + 1) : 0;
        for (u32 kt = 0; kt < KT; kt++) {
            const u32 s = kt % TC_STAGES, ph = (kt / TC_STAGES) & 1;
            const u64 c0 = w0, c1 = w1;
            if (kt + 1 < KT) {
                w0 = row_ok ? __ldg(cp + 2 * (size_t)(kt + 1)) : 0;
                w1 = row_ok ? __ldg(cp + 2 * (size_t)(kt + 1) + 1) : 0;
            }
            tc_mbar_wait(empty(s), ph ^ 1);
            unsigned char *A = smem + (size_t)s * TC_STAGE + (size_t)half * TC_A_HALF + (size_t)row * 128;
#pragma unroll
            for (u32 c = 0; c < 8; c++) {
                const u64 wd = c < 4 ? c0 : c1;
                const u32 h16 = (u32)(wd >> ((c & 3) * 16)) & 0xffffu;
                uint4 v;
                v.x = tc_spread4(h16 & 15u);
                v.y = tc_spread4((h16 >> 4) & 15u);
                v.z = tc_spread4((h16 >> 8) & 15u);
                v.w = tc_spread4((h16 >> 12) & 15u);
                *reinterpret_cast<uint4 *>(A + ((c ^ (row & 7u)) << 4)) = v;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA
            tc_mbar_arrive(full_a(s));
        }
        // ----- epilogue: TMEM lane quarter of this warp is fixed by warp % 4 -----
        tc_mbar_wait(acc_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const u32 lq = warp & 3, ah = (warp - 2) >> 2;
        const size_t b = b0 + ah * 128 + lq * 32 + lane;
        const u32 taddr = tmem_base + ah * 256 + ((lq * 32) << 16);
#pragma unroll 1
        for (u32 cc = 0; cc < 8; cc++) {  // 32 columns = 4 key words per step
            u32 v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr + cc * 32));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (u32 j = 0; j < 4; j++) {
                u64 sum = 0;
#pragma unroll
                for (u32 p = 0; p < 8; p++) sum += (u64)v[j * 8 + p] << (8 * p);
                const u32 x = nt * 32 + cc * 4 + j;
                if (b < batch && x < w) {
                    const u64 lhs = x == kn_out ? ct[b * (size_t)(kn_in + 1) + kn_in] : 0;
                    out[b * (size_t)w + x] = lhs - sum;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

int key_switch_tc_device(const Ksk &k, const u64 *ct, u64 *out, size_t batch, cudaStream_t st) {
    static unsigned long long done_mask = 0;
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    if (!((done_mask >> (dev & 63)) & 1ull)) {
        FHE_CUDA_OK(cudaFuncSetAttribute(ks_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        done_mask |= 1ull << (dev & 63);
    }
    dim3 grid((unsigned)((batch + TC_BM - 1) / TC_BM), k.mma_n_tiles);
    ks_tc_kernel<<<grid, TC_THREADS, TC_SMEM, st>>>(k.mma_blocks, ct, out, batch, (u32)k.kn_in, (u32)k.kn_out);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace fhe
