// microbench.cu -- measures the integer-multiply peaks the NTT rooflines are quoted against (north_star: "as a
// fraction of the slower of the integer-mulmod and HBM rooflines").  Register-only kernels, no memory traffic:
//   kind 0: 32-bit IMAD (mad.lo.u32) lane-ops/s
//   kind 1: 32-bit Shoup modmul/s (the 3-multiply twiddle product of Lazy32/Small32 butterflies)
//   kind 2: 64-bit Shoup modmul/s (Lazy64: one 64x64 high product + two low products)
#include "../../include/fhe_b200.h"
#include "runtime.cuh"

namespace fhe {

template <int KIND>
__global__ void __launch_bounds__(256) int_peak_kernel(u64 *sink, u32 iters, u64 q, u64 w, u64 wp) {
    constexpr int CH = 8;  // independent dependency chains per thread
    const u64 seed = blockIdx.x * 256ull + threadIdx.x + 1;
    if (KIND == 2) {
        Lazy64 m;
        m.q = q; m.q2 = 2 * q; m.qinv_neg = 0; m.qinv = 0; m.r2 = 0;
        const Tw64 t = {w, wp};
        u64 x[CH];
#pragma unroll
        for (int c = 0; c < CH; c++) x[c] = seed * (c + 3);
        for (u32 i = 0; i < iters; i++) {
#pragma unroll
            for (int c = 0; c < CH; c++) x[c] = m.mul_tw(x[c], t);
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
            for (int c = 0; c < CH; c++) x[c] = KIND == 0 ? x[c] * t.w + t.wp : m.mul_tw(x[c], t);
        }
        u32 s = 0;
#pragma unroll
        for (int c = 0; c < CH; c++) s += x[c];
        if (s == 0x1234567u) *sink = s;
    }
}

}  // namespace fhe

using namespace fhe;
extern "C" int fhe_int_peak(int kind, double *ops_per_s) {
    FHE_REQUIRE(ops_per_s != nullptr && kind >= 0 && kind <= 2, "fhe_int_peak: kind must be 0, 1 or 2");
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
    for (int rep = 0; rep < 4; rep++) {
        FHE_CUDA_OK(cudaEventRecord(e0, st));
        if (kind == 0) int_peak_kernel<0><<<grid, 256, 0, st>>>(sink, iters, q32, 12345, 6789);
        else if (kind == 1) int_peak_kernel<1><<<grid, 256, 0, st>>>(sink, iters, q32, 12345, (u64)((12345ull << 32) / q32));
        else int_peak_kernel<2><<<grid, 256, 0, st>>>(sink, iters, q64, 12345, (u64)(((unsigned __int128)12345 << 64) / q64));
        count_launch(1);
        FHE_CUDA_OK(cudaEventRecord(e1, st));
        FHE_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0;
        FHE_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ops_per_s = (double)grid * 256.0 * 8.0 * iters / (best * 1e-3);
    return 0;
}
