// tc_common.cuh -- tcgen05 / mbarrier / bulk-copy PTX wrappers shared by the tensor-core key switch (ks_tc.cu) and
// the int8 tensor-peak microbenchmark (microbench.cu).  sm_100a only.
#pragma once
#include "common.cuh"

namespace fhe {

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

// --- thread-block clusters: one key block fetched once and written into the shared memory of every CTA of the cluster
__device__ __forceinline__ u32 tc_cluster_rank() {
    u32 r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tc_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// `bytes` from global memory to the same shared-memory offset of every CTA in `mask`; each destination CTA's mbarrier at
// the same offset receives the complete_tx
__device__ __forceinline__ void tc_bulk_g2s_multicast(u32 dst, const void *src, u32 bytes, u32 bar, unsigned short mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "h"(mask)
        : "memory");
}
// arrives on the mbarrier at this offset in every CTA of `mask` when the MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_multicast(u32 bar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}


// --- CTA pairs (cta_group::2): one MMA of M = 256 spans the tensor cores of both CTAs of a cluster of two.  CTA r
// supplies rows [128 r, 128 r + 128) of A and rows [128 r, 128 r + 128) of B (N-major) from the SAME shared-memory
// offsets; its tensor memory receives its 128 rows of D.  Only the leader (rank 0) issues.
constexpr u32 TC_IDESC_PAIR = (2u << 4) | (0u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
__device__ __forceinline__ void tc_mma_i8_pair(u32 tmem_d, u64 adesc, u64 bdesc, u32 accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(TC_IDESC_PAIR), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier at this offset in both CTAs of the pair when the pair MMAs issued so far have completed
__device__ __forceinline__ void tc_commit_pair(u32 bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((unsigned short)3)
                 : "memory");
}
// arrive on the mbarrier at the same offset in CTA `rank` of the cluster (release at cluster scope)
__device__ __forceinline__ void tc_mbar_arrive_remote(u32 bar, u32 rank) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(bar), "r"(rank)
        : "memory");
}
// wait with acquire at cluster scope (the arrival came from the peer CTA)
__device__ __forceinline__ void tc_mbar_wait_cluster(u32 bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TC_WAITC:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TC_DONEC;\n"
        "bra TC_WAITC;\n"
        "TC_DONEC:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}

}  // namespace fhe
