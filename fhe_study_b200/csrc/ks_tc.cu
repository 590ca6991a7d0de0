// ks_tc.cu -- key switch on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM).
//
// Same exact byte-plane GEMM as ks_mma.cu (S[b][c] = sum_r A[b][r] * B[r][c], A = bits of the TLWE mask,
// B = u8 planes of the key, s32 accumulation, recombined mod 2^64), same pre-swizzled 32 KB key blocks
// (they are K-major SWIZZLE_128B tiles: 256 rows x 128 B, 8-row groups 1024 B apart).  What changes is the
// engine: the legacy mma.sync path saturates its IMMA pipe at ~0.8 POPS (profiles/r1_keyswitch_imma_*),
// tcgen05 reads both operands from shared memory through descriptors and accumulates in tensor memory.
//
// CTA = 256 ciphertexts x 256 columns (32 key words x 8 planes), K step 128.  Two rings of different depth: the A
// tiles (expanded mask bits, 32 KB per K step) are refilled by the CTA's own threads within a few hundred cycles of
// their stage being freed, the key blocks (32 KB per K step) come from L2 / HBM with a round trip of about two K steps
// of MMA time -- so A gets 3 stages and B gets 4 (224 KB of shared memory in all; r1 ran 3 + 3 and the tensor pipe
// idled 18 % of the time waiting for key blocks):
//   warp 0      : one lane streams the key blocks with cp.async.bulk (TMA 1-D) -> full_b[s]
//   warp 1      : allocates TMEM (512 columns = two 128x256 s32 accumulators); one lane issues, per K step,
//                 4 x 2 tcgen05.mma (M=128, N=256, K=32) and commits them to empty_a[sa] and empty_b[sb]
//   warps 2..9  : 256 threads, one ciphertext row each: expand 2 mask words per K step into 128 bytes of the
//                 swizzled A tiles (never materialised in HBM), fence.proxy.async, arrive on full_a[s];
//                 at the end they are the epilogue: tcgen05.ld the accumulator rows, shift-and-add the 8 plane
//                 sums of each key word, subtract from (0,..,0,b) (tfhe/src/tlwe.rs:111) and store.
#include "../../include/fhe_b200.h"
#include "runtime.cuh"
#include <stdlib.h>

#include <atomic>

#include "tc_common.cuh"
#include "tlwe.cuh"

namespace fhe {

constexpr int TC_BM = 256, TC_BN = 256, TC_BK = 128;
#ifndef FHE_KS_DEFAULT_MODE
#define FHE_KS_DEFAULT_MODE 2
#endif
#ifndef FHE_KS_STAGES_A
#define FHE_KS_STAGES_A 3
#endif
#ifndef FHE_KS_STAGES_B
#define FHE_KS_STAGES_B 3
#endif
constexpr int TC_SA = FHE_KS_STAGES_A, TC_SB = FHE_KS_STAGES_B;
constexpr int TC_THREADS = 320;                       // 10 warps
constexpr int TC_A_HALF = 128 * TC_BK;                // 16 KB: one 128-row A tile
constexpr int TC_A_BYTES = 2 * TC_A_HALF;             // 32 KB: both A tiles of a K step
constexpr int TC_B_BYTES = TC_BN * TC_BK;             // 32 KB
constexpr size_t TC_RING = (size_t)TC_SA * TC_A_BYTES + (size_t)TC_SB * TC_B_BYTES;
constexpr size_t TC_SMEM = TC_RING + 256;
static_assert(TC_SMEM <= 227 * 1024, "ks_tc: rings exceed one CTA's shared memory");
static_assert((2 * TC_SA + 2 * TC_SB + 1) * 8 + 4 <= 256, "ks_tc: barrier block too small");

__device__ __forceinline__ u32 tc_spread4(u32 nib) { return (nib * 0x00204081u) & 0x01010101u; }

// CL = CTAs per cluster (1 or 2).  With CL = 2 the two CTAs of a cluster (neighbouring blocks of 256 ciphertexts, same
// column tile) need the same key blocks: each fetches HALF of every block and multicasts it into both shared memories,
// so the L2 -> SM key traffic halves.  At 3 M bootstraps/s that traffic is 8.6 TB/s for CL = 1 -- the chip-wide L2
// limit (~12 TB/s at full clock) minus what the mask words take -- which is what held the tensor pipe at 82 %.
template <int CL>
__global__ void __launch_bounds__(TC_THREADS, 1)
ks_tc_kernel(const unsigned char *__restrict__ blocks, const u64 *__restrict__ ct, u64 *__restrict__ out, size_t batch,
             u32 kn_in, u32 kn_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *ringA = smem, *ringB = smem + (size_t)TC_SA * TC_A_BYTES;
    u64 *bars = reinterpret_cast<u64 *>(smem + TC_RING);
    // bars: [0,SA) full_a, [SA,2SA) empty_a, [2SA,2SA+SB) full_b, [2SA+SB,2SA+2SB) empty_b, then accumulators ready
    u32 *tmem_slot = reinterpret_cast<u32 *>(bars + 2 * TC_SA + 2 * TC_SB + 1);
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u32 KT = kn_in / 2, w = kn_out + 1;
    const size_t b0 = (size_t)blockIdx.x * TC_BM;
    const u32 nt = blockIdx.y;
    const unsigned char *gB = blocks + (size_t)nt * KT * TC_B_BYTES;
    auto full_a = [&](u32 s) { return tc_smem_u32(&bars[s]); };
    auto empty_a = [&](u32 s) { return tc_smem_u32(&bars[TC_SA + s]); };
    auto full_b = [&](u32 s) { return tc_smem_u32(&bars[2 * TC_SA + s]); };
    auto empty_b = [&](u32 s) { return tc_smem_u32(&bars[2 * TC_SA + TC_SB + s]); };
    const u32 acc_bar = tc_smem_u32(&bars[2 * TC_SA + 2 * TC_SB]);

    if (tid == 0) {
        for (u32 s = 0; s < TC_SA; s++) {
            tc_mbar_init(full_a(s), 256);
            tc_mbar_init(empty_a(s), 1);
        }
        for (u32 s = 0; s < TC_SB; s++) {
            tc_mbar_init(full_b(s), 1);
            tc_mbar_init(empty_b(s), CL);  // the stage is rewritten for every CTA of the cluster at once
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
    if constexpr (CL > 1) tc_cluster_sync();  // the peer's barriers are initialised before anything arrives on them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem_base = *tmem_slot;
    constexpr unsigned short CL_MASK = (unsigned short)((1u << CL) - 1u);

    if (warp == 0) {
        // ===== key-block producer =====
        if (lane == 0) {
            const u32 rank = CL > 1 ? tc_cluster_rank() : 0u;
            constexpr u32 PART = TC_B_BYTES / CL;
            for (u32 kt = 0; kt < KT; kt++) {
                const u32 s = kt % TC_SB, ph = (kt / TC_SB) & 1;
                tc_mbar_wait(empty_b(s), ph ^ 1);  // stage s is free in EVERY CTA of the cluster
                tc_mbar_expect_tx(full_b(s), TC_B_BYTES);
                const u32 dst = tc_smem_u32(ringB + (size_t)s * TC_B_BYTES + (size_t)rank * PART);
                const unsigned char *src = gB + (size_t)kt * TC_B_BYTES + (size_t)rank * PART;
                if constexpr (CL > 1) tc_bulk_g2s_multicast(dst, src, PART, full_b(s), CL_MASK);
                else tc_bulk_g2s(dst, src, PART, full_b(s));
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            for (u32 kt = 0; kt < KT; kt++) {
                const u32 sa_i = kt % TC_SA, pha = (kt / TC_SA) & 1, sb_i = kt % TC_SB, phb = (kt / TC_SB) & 1;
                tc_mbar_wait(full_a(sa_i), pha);
                tc_mbar_wait(full_b(sb_i), phb);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const u32 sa = tc_smem_u32(ringA + (size_t)sa_i * TC_A_BYTES), sb = tc_smem_u32(ringB + (size_t)sb_i * TC_B_BYTES);
#pragma unroll
                for (u32 k = 0; k < TC_BK / 32; k++) {
                    const u64 bd = tc_smem_desc(sb + k * 32);
                    const u32 accum = (kt | k) != 0 ? 1u : 0u;
                    tc_mma_i8(tmem_base, tc_smem_desc(sa + k * 32), bd, accum);
                    tc_mma_i8(tmem_base + 256, tc_smem_desc(sa + TC_A_HALF + k * 32), bd, accum);
                }
                tc_commit(empty_a(sa_i));  // both arrive when the MMAs above have finished reading their stages
                if constexpr (CL > 1) tc_commit_multicast(empty_b(sb_i), CL_MASK);
                else tc_commit(empty_b(sb_i));
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
            const u32 s = kt % TC_SA, ph = (kt / TC_SA) & 1;
            const u64 c0 = w0, c1 = w1;
            if (kt + 1 < KT) {
                w0 = row_ok ? __ldg(cp + 2 * (size_t)(kt + 1)) : 0;
                w1 = row_ok ? __ldg(cp + 2 * (size_t)(kt + 1) + 1) : 0;
            }
            tc_mbar_wait(empty_a(s), ph ^ 1);
            unsigned char *A = ringA + (size_t)s * TC_A_BYTES + (size_t)half * TC_A_HALF + (size_t)row * 128;
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
    if constexpr (CL > 1) tc_cluster_sync();  // no CTA leaves while its peer may still write into it or arrive on its barriers
}

// FHE_KS_CLUSTER=1|2|3: single CTAs, clusters of two with multicast key blocks, CTA pairs (cta_group::2 MMAs)
static int ks_cluster_size() {
    static const int v = [] {
        const char *e = getenv("FHE_KS_CLUSTER");
        const int c = e ? atoi(e) : FHE_KS_DEFAULT_MODE;
        return c == 1 ? 1 : c == 3 ? 3 : 2;
    }();
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2).  Same tile per CTA (256 ciphertexts x 256 columns, two accumulators in TMEM), but each
// MMA is M = 256: the upper 128 rows come from the peer CTA's A tile, and each CTA holds only ITS half of the key
// block (128 of the 256 N-rows: 16 KB per K step instead of 32).  Per K step an SM's tensor core therefore reads
// 32 KB of A + 2 x 16 KB of B instead of 32 + 2 x 32 KB, and the key ring shrinks by half -- room for a deeper one with
// L1 left for the mask words.  The leader's MMA thread needs the PEER's stages too: the peer's otherwise idle MMA warp
// forwards "my A and B stages of step kt are full" to the leader's peer_ready ring.
// Ring depths are launch parameters (FHE_KS_PAIR_SA / FHE_KS_PAIR_SB override them for tuning).
#ifndef FHE_KS_PAIR_STAGES_A
#define FHE_KS_PAIR_STAGES_A 4
#endif
#ifndef FHE_KS_PAIR_STAGES_B
#define FHE_KS_PAIR_STAGES_B 4
#endif
constexpr int TP_MAX_SA = 6, TP_MAX_SB = 8, TP_PR = 8;
constexpr int TP_B_BYTES = TC_B_BYTES / 2;
constexpr int TP_BAR_BYTES = 512;
static_assert((2 * TP_MAX_SA + 2 * TP_MAX_SB + 1 + TP_PR) * 8 + 4 <= TP_BAR_BYTES, "ks_tc pair: barrier block too small");
static_assert(TP_MAX_SB <= TP_PR && TP_MAX_SA < TP_PR, "peer_ready slots are reused only behind the rings");
struct PairRing {
    u32 i = 0, ph = 0, n;
    __device__ explicit PairRing(u32 n_) : n(n_) {}
    __device__ void next() { if (++i == n) { i = 0; ph ^= 1; } }
};

__global__ void __launch_bounds__(TC_THREADS, 1)
ks_tc_pair_kernel(const unsigned char *__restrict__ blocks, const u64 *__restrict__ ct, u64 *__restrict__ out, size_t batch,
                  u32 kn_in, u32 kn_out, u32 TP_SA, u32 TP_SB) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *ringA = smem, *ringB = smem + (size_t)TP_SA * TC_A_BYTES;
    u64 *bars = reinterpret_cast<u64 *>(ringB + (size_t)TP_SB * TP_B_BYTES);
    // bars: full_a[SA], empty_a[SA], full_b[SB], empty_b[SB], accumulators ready, peer_ready[PR]
    u32 *tmem_slot = reinterpret_cast<u32 *>(bars + 2 * TP_SA + 2 * TP_SB + 1 + TP_PR);
    const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const u32 KT = kn_in / 2, w = kn_out + 1;
    const size_t b0 = (size_t)blockIdx.x * TC_BM;
    const u32 nt = blockIdx.y;
    const u32 rank = tc_cluster_rank();
    const unsigned char *gB = blocks + (size_t)nt * KT * TC_B_BYTES + (size_t)rank * TP_B_BYTES;
    auto full_a = [&](u32 s) { return tc_smem_u32(&bars[s]); };
    auto empty_a = [&](u32 s) { return tc_smem_u32(&bars[TP_SA + s]); };
    auto full_b = [&](u32 s) { return tc_smem_u32(&bars[2 * TP_SA + s]); };
    auto empty_b = [&](u32 s) { return tc_smem_u32(&bars[2 * TP_SA + TP_SB + s]); };
    const u32 acc_bar = tc_smem_u32(&bars[2 * TP_SA + 2 * TP_SB]);
    auto peer_ready = [&](u32 s) { return tc_smem_u32(&bars[2 * TP_SA + 2 * TP_SB + 1 + s]); };

    if (tid == 0) {
        for (u32 s = 0; s < TP_SA; s++) {
            tc_mbar_init(full_a(s), 256);
            tc_mbar_init(empty_a(s), 1);
        }
        for (u32 s = 0; s < TP_SB; s++) {
            tc_mbar_init(full_b(s), 1);
            tc_mbar_init(empty_b(s), 1);
        }
        tc_mbar_init(acc_bar, 1);
        for (u32 s = 0; s < TP_PR; s++) tc_mbar_init(peer_ready(s), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM of the pair: one warp of each CTA, all 512 columns
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    tc_cluster_sync();  // both CTAs' barriers and tensor memory exist before anything arrives or is issued
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== key-block producer: this CTA's half of every block =====
        if (lane == 0) {
            PairRing rb(TP_SB);
            for (u32 kt = 0; kt < KT; kt++, rb.next()) {
                tc_mbar_wait(empty_b(rb.i), rb.ph ^ 1);
                tc_mbar_expect_tx(full_b(rb.i), TP_B_BYTES);
                tc_bulk_g2s(tc_smem_u32(ringB + (size_t)rb.i * TP_B_BYTES), gB + (size_t)kt * TC_B_BYTES, TP_B_BYTES, full_b(rb.i));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            if (rank == 0) {
                // ===== MMA issuer of the pair =====
                PairRing ra(TP_SA), rb(TP_SB);
                for (u32 kt = 0; kt < KT; kt++, ra.next(), rb.next()) {
                    const u32 sa_i = ra.i, sb_i = rb.i;
                    tc_mbar_wait(full_a(sa_i), ra.ph);
                    tc_mbar_wait(full_b(sb_i), rb.ph);
                    tc_mbar_wait_cluster(peer_ready(kt % TP_PR), (kt / TP_PR) & 1);   // the peer's A tiles and key half
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const u32 sa = tc_smem_u32(ringA + (size_t)sa_i * TC_A_BYTES), sb = tc_smem_u32(ringB + (size_t)sb_i * TP_B_BYTES);
#pragma unroll
                    for (u32 k = 0; k < TC_BK / 32; k++) {
                        const u64 bd = tc_smem_desc(sb + k * 32);
                        const u32 accum = (kt | k) != 0 ? 1u : 0u;
                        tc_mma_i8_pair(tmem_base, tc_smem_desc(sa + k * 32), bd, accum);
                        tc_mma_i8_pair(tmem_base + 256, tc_smem_desc(sa + TC_A_HALF + k * 32), bd, accum);
                    }
                    tc_commit_pair(empty_a(sa_i));  // both CTAs' stages are free when these MMAs have read them
                    tc_commit_pair(empty_b(sb_i));
                }
                tc_commit_pair(acc_bar);
            } else {
                // ===== peer: tells the leader when this CTA's stages of step kt are full.  (Measured: the expander warps
                // arriving on the leader's barrier themselves -- eight release.cluster arrivals per stage -- 1.94 against
                // 2.84 M bootstraps/s with this single forwarding thread.)
                PairRing ra(TP_SA), rb(TP_SB);
                for (u32 kt = 0; kt < KT; kt++, ra.next(), rb.next()) {
                    tc_mbar_wait(full_a(ra.i), ra.ph);
                    tc_mbar_wait(full_b(rb.i), rb.ph);
                    tc_mbar_arrive_remote(peer_ready(kt % TP_PR), 0);
                }
            }
        }
    } else {
        // ===== A expanders (one ciphertext row per thread), then the epilogue: as in ks_tc_kernel =====
        const u32 e = tid - 64;  // 0..255
        const u32 half = e >> 7, row = e & 127;
        const bool row_ok = (b0 + e) < batch;
        const u64 *cp = ct + (b0 + e) * (size_t)(kn_in + 1);
        u64 w0 = row_ok ? __ldg(cp) : 0, w1 = row_ok ? __ldg(cp + 1) : 0;
        PairRing ra(TP_SA);
        for (u32 kt = 0; kt < KT; kt++, ra.next()) {
            const u32 s = ra.i, ph = ra.ph;
            const u64 c0 = w0, c1 = w1;
            if (kt + 1 < KT) {
                w0 = row_ok ? __ldg(cp + 2 * (size_t)(kt + 1)) : 0;
                w1 = row_ok ? __ldg(cp + 2 * (size_t)(kt + 1) + 1) : 0;
            }
            tc_mbar_wait(empty_a(s), ph ^ 1);
            unsigned char *A = ringA + (size_t)s * TC_A_BYTES + (size_t)half * TC_A_HALF + (size_t)row * 128;
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
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_mbar_arrive(full_a(s));
        }
        tc_mbar_wait(acc_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const u32 lq = warp & 3, ah = (warp - 2) >> 2;
        const size_t b = b0 + ah * 128 + lq * 32 + lane;
        const u32 taddr = tmem_base + ah * 256 + ((lq * 32) << 16);
#pragma unroll 1
        for (u32 cc = 0; cc < 8; cc++) {
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
    tc_cluster_sync();  // the pair's MMAs read both shared memories and write both tensor memories: leave together
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

static int launch_ks_tc_pair(const Ksk &k, const u64 *ct, u64 *out, size_t batch, cudaStream_t st) {
    static std::atomic<unsigned long long> done_mask{0};
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    if (!((done_mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
        FHE_CUDA_OK(cudaFuncSetAttribute(ks_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        done_mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    static const u32 sa = [] { const char *e = getenv("FHE_KS_PAIR_SA"); const int v = e ? atoi(e) : FHE_KS_PAIR_STAGES_A; return (u32)(v < 2 ? 2 : v > TP_MAX_SA ? TP_MAX_SA : v); }();
    static const u32 sb = [] { const char *e = getenv("FHE_KS_PAIR_SB"); const int v = e ? atoi(e) : FHE_KS_PAIR_STAGES_B; return (u32)(v < 2 ? 2 : v > TP_MAX_SB ? TP_MAX_SB : v); }();
    const size_t TP_SMEM = (size_t)sa * TC_A_BYTES + (size_t)sb * TP_B_BYTES + TP_BAR_BYTES;
    FHE_REQUIRE(TP_SMEM <= 227 * 1024, "key switch (CTA pairs): rings exceed one CTA's shared memory");
    unsigned gx = (unsigned)((batch + TC_BM - 1) / TC_BM);
    gx = (gx + 1) / 2 * 2;  // whole pairs; a block past the end of the batch only feeds its peer
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, k.mma_n_tiles);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TP_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FHE_CUDA_OK(cudaLaunchKernelEx(&cfg, ks_tc_pair_kernel, (const unsigned char *)k.mma_blocks, ct, out, batch, (u32)k.kn_in,
                                   (u32)k.kn_out, sa, sb));
    count_launch(1);
    return 0;
}

template <int CL>
static int launch_ks_tc(const Ksk &k, const u64 *ct, u64 *out, size_t batch, cudaStream_t st) {
    static std::atomic<unsigned long long> done_mask{0};
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    if (!((done_mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
        FHE_CUDA_OK(cudaFuncSetAttribute(ks_tc_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM));
        done_mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    unsigned gx = (unsigned)((batch + TC_BM - 1) / TC_BM);
    gx = (gx + CL - 1) / CL * CL;  // whole clusters; a block past the end of the batch only feeds its peer
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, k.mma_n_tiles);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = TC_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FHE_CUDA_OK(cudaLaunchKernelEx(&cfg, ks_tc_kernel<CL>, (const unsigned char *)k.mma_blocks, ct, out, batch, (u32)k.kn_in,
                                   (u32)k.kn_out));
    count_launch(1);
    return 0;
}
int key_switch_tc_device(const Ksk &k, const u64 *ct, u64 *out, size_t batch, cudaStream_t st) {
    const int mode = ks_cluster_size();
    if (mode == 3) return launch_ks_tc_pair(k, ct, out, batch, st);
    return mode == 2 ? launch_ks_tc<2>(k, ct, out, batch, st) : launch_ks_tc<1>(k, ct, out, batch, st);
}

}  // namespace fhe
