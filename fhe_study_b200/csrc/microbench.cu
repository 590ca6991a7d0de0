// microbench.cu -- measures the integer-multiply peaks the NTT rooflines are quoted against (north_star: "as a
// fraction of the slower of the integer-mulmod and HBM rooflines").  Register-only kernels, no memory traffic:
//   kind 0: 32-bit IMAD (mad.lo.u32) lane-ops/s
//   kind 1: 32-bit Shoup modmul/s (the 3-multiply twiddle product of Lazy32/Small32 butterflies)
//   kind 2: 64-bit Shoup modmul/s (Lazy64: one 64x64 high product + two low products) -- the better of the library's
//           chained form (shoup_tail64) and the plain  y*w - mulhi(y,w')*q  the compiler schedules itself, so that a change
//           of the library's own formulation cannot lower the denominator its kernels are quoted against
//   kind 6: the same 32-bit product in the Fermat form available for q = 2^16 + 1 (IMAD.WIDE + shift + IMAD + add): an
//           experiment's evidence, not a roofline denominator
//   kind 3: int8 tensor-core ops/s (2 per MAC): back-to-back tcgen05.mma.kind::i8 (M=128, N=256, K=32) on operands
//           resident in shared memory, no loads, no epilogue -- the measured peak the key-switch GEMM is quoted against
#include "../../include/fhe_b200.h"
#include "runtime.cuh"
#include "tc_common.cuh"

#include <atomic>

namespace fhe {

template <int KIND>
__global__ void __launch_bounds__(256) int_peak_kernel(u64 *sink, u32 iters, u64 q, u64 w, u64 wp) {
    constexpr int CH = 8;  // independent dependency chains per thread
    const u64 seed = blockIdx.x * 256ull + threadIdx.x + 1;
    if (KIND == 2 || KIND == 5) {
        Lazy64 m;
        m.q = q; m.q2 = 2 * q; m.qinv_neg = 0; m.qinv = 0; m.r2 = 0; m.nq = (u64)0 - q;
        const Tw64 t = {w, wp};
        u64 x[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) x[c] = seed * (c + 3);
        for (u32 i = 0; i < iters; i++) {
#pragma unroll
            for (int c = 0; c < CH; c++) x[c] = KIND == 2 ? m.mul_tw(x[c], t) : x[c] * t.w - mulhi_u64(x[c], t.wp) * q;
        }
        u64 s = 0;
#pragma unroll
        for (int c = 0; c < CH; c++) s += x[c];
        if (s == 0x1234567ull) *sink = s;
    } else {
        Lazy32 m;
        m.q = (u32)q; m.q2 = 2 * (u32)q; m.bk_shift = 0; m.bk_mu = 0;
        const Tw32 t = {(u32)w, (u32)wp};
        u32 x[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) x[c] = (u32)seed * (c + 3);
        for (u32 i = 0; i < iters; i++) {
#pragma unroll
            for (int c = 0; c < CH; c++) {
                if (KIND == 6) {   // Fermat-prime product (q = 2^16 + 1): y*w = hi*2^32 + lo == lo - (lo >> 16)*q + hi
                    const u64 T = (u64)x[c] * t.w;
                    const u32 lo = (u32)T, hi = (u32)(T >> 32);
                    x[c] = (lo >> 16) * t.wp + lo + (hi + m.q2);   // t.wp = 2^32 - q
                } else {
                    x[c] = KIND == 0 ? x[c] * t.w + t.wp : m.mul_tw(x[c], t);
                }
            }
        }
        u32 s = 0;
#pragma unroll
        for (int c = 0; c < CH; c++) s += x[c];
        if (s == 0x1234567u) *sink = s;
    }
}

// One CTA per SM.  Shared memory holds one 128x128-byte A tile and one 256x128-byte B tile (whatever bytes happen to be
// there: integer MACs have no special values); one thread issues `iters` x 8 MMAs alternating between the two TMEM
// accumulators (the key switch's own issue pattern, ks_tc.cu), commits, and waits for the last one to retire.
constexpr int PK_SMEM = 16 * 1024 + 32 * 1024 + 1024 + 64;
__global__ void __launch_bounds__(64, 1) int8_peak_kernel(u32 iters) {
    extern __shared__ __align__(1024) unsigned char smem_pk[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_pk) + 1023) & ~(uintptr_t)1023);
    u64 *bar = reinterpret_cast<u64 *>(base + 48 * 1024);
    u32 *tmem_slot = reinterpret_cast<u32 *>(bar + 1);
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (u32 i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<u32 *>(base)[i] = i * 2654435761u;
    if (threadIdx.x == 0) {
        tc_mbar_init(tc_smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 tmem_base = *tmem_slot;
    if (warp == 0 && lane == 0) {
        const u32 sa = tc_smem_u32(base), sb = sa + 16 * 1024;
        for (u32 it = 0; it < iters; it++) {
#pragma unroll
            for (u32 k = 0; k < 4; k++) {
                const u64 bd = tc_smem_desc(sb + k * 32), ad = tc_smem_desc(sa + k * 32);
                tc_mma_i8(tmem_base, ad, bd, (it | k) != 0 ? 1u : 0u);
                tc_mma_i8(tmem_base + 256, ad, bd, (it | k) != 0 ? 1u : 0u);
            }
        }
        tc_commit(tc_smem_u32(bar));
        tc_mbar_wait(tc_smem_u32(bar), 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace fhe

using namespace fhe;
static int int8_tensor_peak(double *ops_per_s, double seconds) {
    cudaStream_t st = current_stream();
    static std::atomic<unsigned long long> done_mask{0};
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    if (!((done_mask.load() >> (dev & 63)) & 1ull)) {
        FHE_CUDA_OK(cudaFuncSetAttribute(int8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM));
        done_mask.fetch_or(1ull << (dev & 63));
    }
    const unsigned grid = (unsigned)num_sms();
    const u32 iters = 8192;  // x 8 MMAs of 2^20 MACs: ~0.3 ms per launch at the nominal rate
    cudaEvent_t e0, e1;
    FHE_CUDA_OK(cudaEventCreate(&e0));
    FHE_CUDA_OK(cudaEventCreate(&e1));
    const double per_launch = (double)grid * iters * 8.0 * 2.0 * 128.0 * 256.0 * 32.0;
    double best = 0;
    int rc = 0;
    if (seconds <= 0) {  // burst: best of a few launches
        for (int rep = 0; rep < 5 && !rc; rep++) {
            cudaEventRecord(e0, st);
            int8_peak_kernel<<<grid, 64, PK_SMEM, st>>>(iters);
            count_launch(1);
            cudaEventRecord(e1, st);
            if (cudaEventSynchronize(e1) != cudaSuccess) { rc = -2; break; }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && per_launch / (ms * 1e-3) > best) best = per_launch / (ms * 1e-3);
        }
    } else {  // sustained: launches back to back for `seconds`, one rate over the whole interval
        const int chunk = 64;
        double elapsed = 0, launches = 0;
        while (elapsed < seconds && !rc) {
            cudaEventRecord(e0, st);
            for (int i = 0; i < chunk; i++) int8_peak_kernel<<<grid, 64, PK_SMEM, st>>>(iters);
            count_launch(chunk);
            cudaEventRecord(e1, st);
            if (cudaEventSynchronize(e1) != cudaSuccess) { rc = -2; break; }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            elapsed += ms * 1e-3;
            launches += chunk;
        }
        if (!rc) best = per_launch * launches / elapsed;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) {
        set_error(std::string("int8 tensor peak: ") + cudaGetErrorString(cudaGetLastError()));
        return rc;
    }
    *ops_per_s = best;
    return 0;
}
// kind 3 = burst, kind 4 = sustained over ~2 s (the clock settles under the power cap)
extern "C" int fhe_int_peak(int kind, double *ops_per_s) {
    FHE_REQUIRE(ops_per_s != nullptr && kind >= 0 && kind <= 6 && kind != 5, "fhe_int_peak: kind must be 0..4 or 6");
    if (kind == 3 || kind == 4) return int8_tensor_peak(ops_per_s, kind == 4 ? 2.0 : 0.0);
    cudaStream_t st = current_stream();
    Scratch s_sink;
    int rc0 = s_sink.alloc(8, st);
    if (rc0) return rc0;
    u64 *sink = s_sink.ptr<u64>();
    const u32 iters = 4096;
    const unsigned grid = (unsigned)num_sms() * 16;
    const u64 q32 = 0x7E90001ull, q64 = 0x3FFFFFFFFFFF0001ull;
    cudaEvent_t e0, e1;
    FHE_CUDA_OK(cudaEventCreate(&e0));
    FHE_CUDA_OK(cudaEventCreate(&e1));
    float best = 1e30f;
    const u64 wp64 = (u64)(((unsigned __int128)12345 << 64) / q64);
    for (int rep = 0; rep < (kind == 2 ? 8 : 4); rep++) {
        FHE_CUDA_OK(cudaEventRecord(e0, st));
        if (kind == 0) int_peak_kernel<0><<<grid, 256, 0, st>>>(sink, iters, q32, 12345, 6789);
        else if (kind == 1) int_peak_kernel<1><<<grid, 256, 0, st>>>(sink, iters, q32, 12345, (u64)((12345ull << 32) / q32));
        else if (kind == 6) int_peak_kernel<6><<<grid, 256, 0, st>>>(sink, iters, 65537, 12345, (u64)(0u - 65537u));
        else if (rep < 4) int_peak_kernel<2><<<grid, 256, 0, st>>>(sink, iters, q64, 12345, wp64);
        else int_peak_kernel<5><<<grid, 256, 0, st>>>(sink, iters, q64, 12345, wp64);
        count_launch(1);
        FHE_CUDA_OK(cudaEventRecord(e1, st));
        FHE_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0;
        FHE_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        if ((rep & 3) > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ops_per_s = (double)grid * 256.0 * 8.0 * iters / (best * 1e-3);
    return 0;
}
