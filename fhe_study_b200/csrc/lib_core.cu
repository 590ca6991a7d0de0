// lib_core.cu -- runtime state (errors, stream, launch counter) and the NTT-plan / Rq entry points of the
// C ABI declared in include/fhe_b200.h.
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <memory>
#include <tuple>

#include "../../include/fhe_b200.h"
#include "plan.cuh"
#include "runtime.cuh"

namespace fhe {

static thread_local std::string t_error;
static thread_local cudaStream_t t_stream = nullptr;
static std::atomic<unsigned long long> g_launches{0};

void set_error(const std::string &msg) { t_error = msg; }
cudaStream_t current_stream() { return t_stream; }
void count_launch(unsigned long long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// keep stream-ordered scratch cached in the pool instead of returning it to the OS at every synchronisation
void device_init_once() {
    static unsigned long long done_mask = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    if ((done_mask >> (dev & 63)) & 1ull) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
    done_mask |= 1ull << (dev & 63);
}
int num_sms() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (sms[dev & 63] == 0) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        sms[dev & 63] = p.multiProcessorCount;
    }
    return sms[dev & 63];
}

}  // namespace fhe

using namespace fhe;

namespace {
std::mutex g_plan_mu;
std::map<std::tuple<int, u64, u64>, fhe_ntt_plan *> g_plans;

template <class M> int upload_tables(fhe_ntt_plan *p, NttParams<M> &dst) {
    ExpandedTables<M> x;
    expand_tables(p->host, x, p->loge);
    const size_t bytes = sizeof(typename M::T) * p->host.n;
    FHE_CUDA_OK(cudaMalloc(&p->d_fwd, bytes));
    FHE_CUDA_OK(cudaMalloc(&p->d_inv, bytes));
    FHE_CUDA_OK(cudaMemcpy(p->d_fwd, x.fwd.data(), bytes, cudaMemcpyHostToDevice));
    FHE_CUDA_OK(cudaMemcpy(p->d_inv, x.inv.data(), bytes, cudaMemcpyHostToDevice));
    dst.mod = x.mod;
    dst.fwd = reinterpret_cast<const typename M::T *>(p->d_fwd);
    dst.inv = reinterpret_cast<const typename M::T *>(p->d_inv);
    dst.ninv = x.ninv;
    dst.s_ninv = x.s_ninv;
    dst.ninv_pw = x.ninv_pw;
    dst.s_ninv_pw = x.s_ninv_pw;
    for (u64 i = 0; i < 64; i++) {
        dst.c_fwd[i] = x.fwd[i < p->host.n ? i : 0];
        dst.c_inv[i] = x.inv[i < p->host.n ? i : 0];
    }
    return 0;
}

int launch_plan(const fhe_ntt_plan *plan, int mode, const u64 *a, const u64 *b, u64 *c, u64 *c_evals, size_t batch,
                int flags, cudaStream_t st) {
    int rc;
    switch (plan->kind) {
        case 3: rc = ntt_launch_small32(plan->logn, plan->loge, mode, plan->psm, a, b, c, c_evals, batch, flags, st); break;
        case 0: rc = ntt_launch_lazy32(plan->logn, plan->loge, mode, plan->p32, a, b, c, c_evals, batch, flags, st); break;
        case 1: rc = ntt_launch_lazy64(plan->logn, plan->loge, mode, plan->p64, a, b, c, c_evals, batch, flags, st); break;
        default: rc = ntt_launch_strict64(plan->logn, plan->loge, mode, plan->ps64, a, b, c, c_evals, batch, flags, st);
    }
    if (!rc) count_launch(1);
    return rc;
}


thread_local PipeStreams t_pipe;

// bytes per operand and pipeline stage of the chunked host-buffer paths (FHE_PIPE_CHUNK_MB overrides; tuning knob).
// The first H2D and the last D2H of a call are not overlapped with anything, so smaller stages shorten fill and drain.
size_t pipe_chunk_bytes() {
    static const size_t v = [] {
        const char *e = getenv("FHE_PIPE_CHUNK_MB");
        const long mb = e ? atol(e) : 0;
        return (size_t)(mb >= 1 && mb <= 1024 ? mb : 32) << 20;
    }();
    return v;
}

int run_ntt_pipelined(const fhe_ntt_plan *plan, int mode, const u64 *a, const u64 *b, u64 *c, u64 *c_evals, size_t batch,
                      int flags, cudaStream_t st, size_t chunk) {
    int rc = t_pipe.init();
    if (rc) return rc;
    PipeStreams &ps = t_pipe;
    const size_t n = plan->host.n, cbytes = chunk * n * sizeof(u64);
    const int nbuf = 1 + (b ? 1 : 0) + 1 + (c_evals ? 1 : 0);
    Scratch scratch;
    if ((rc = scratch.alloc(2 * nbuf * cbytes, st))) return rc;
    u64 *dev = scratch.ptr<u64>();
    FHE_CUDA_OK(cudaStreamSynchronize(st));  // the scratch is used from the side streams as well
    auto buf = [&](int which, int parity) { return dev + ((size_t)parity * nbuf + which) * chunk * n; };
    const int ib = 1, ic = b ? 2 : 1, ie = ic + 1;
    size_t i = 0;
    for (size_t off = 0; off < batch; off += chunk, i++) {
        const size_t nb = std::min(chunk, batch - off), bytes = nb * n * sizeof(u64);
        const int par = (int)(i & 1);
        if (i >= 2) FHE_CUDA_OK(cudaStreamWaitEvent(ps.h2d, ps.comp_done[par], 0));  // inputs of chunk i-2 consumed
        FHE_CUDA_OK(cudaMemcpyAsync(buf(0, par), a + off * n, bytes, cudaMemcpyHostToDevice, ps.h2d));
        if (b) FHE_CUDA_OK(cudaMemcpyAsync(buf(ib, par), b + off * n, bytes, cudaMemcpyHostToDevice, ps.h2d));
        FHE_CUDA_OK(cudaEventRecord(ps.h2d_done[par], ps.h2d));
        FHE_CUDA_OK(cudaStreamWaitEvent(st, ps.h2d_done[par], 0));
        if (i >= 2) FHE_CUDA_OK(cudaStreamWaitEvent(st, ps.d2h_done[par], 0));       // outputs of chunk i-2 drained
        if ((rc = launch_plan(plan, mode, buf(0, par), b ? buf(ib, par) : nullptr, buf(ic, par),
                              c_evals ? buf(ie, par) : nullptr, nb, flags, st)))
            break;
        FHE_CUDA_OK(cudaEventRecord(ps.comp_done[par], st));
        FHE_CUDA_OK(cudaStreamWaitEvent(ps.d2h, ps.comp_done[par], 0));
        FHE_CUDA_OK(cudaMemcpyAsync(c + off * n, buf(ic, par), bytes, cudaMemcpyDeviceToHost, ps.d2h));
        if (c_evals) FHE_CUDA_OK(cudaMemcpyAsync(c_evals + off * n, buf(ie, par), bytes, cudaMemcpyDeviceToHost, ps.d2h));
        FHE_CUDA_OK(cudaEventRecord(ps.d2h_done[par], ps.d2h));
    }
    cudaError_t e1 = cudaStreamSynchronize(ps.d2h), e2 = cudaStreamSynchronize(ps.h2d), e3 = cudaStreamSynchronize(st);
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        set_error(std::string("pipelined transfer failed: ") +
                  cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
        return -2;
    }
    return 0;
}

int run_ntt(const fhe_ntt_plan *plan, int mode, const u64 *a, const u64 *b, u64 *c, u64 *c_evals, size_t batch,
            int flags) {
    FHE_REQUIRE(plan != nullptr, "null plan");
    if (batch == 0) return 0;
    FHE_REQUIRE(a != nullptr && c != nullptr && (mode != MODE_MUL || b != nullptr), "null polynomial pointer");
    cudaStream_t st = current_stream();
    const size_t bytes = batch * plan->host.n * sizeof(u64);
    if (mode != MODE_MUL) b = nullptr;
    const bool bcast = mode == MODE_MUL && (flags & B_BROADCAST);
    if (!bcast) {   // all-host call on a batch worth pipelining (>= 4 chunks of ~32 MiB per operand)
        const size_t chunk = std::max<size_t>(1, pipe_chunk_bytes() / (plan->host.n * sizeof(u64)));
        if (batch >= 4 * chunk && is_host_ptr(a) && (!b || is_host_ptr(b)) && is_host_ptr(c) &&
            (!c_evals || is_host_ptr(c_evals)) && a != c && b != c)
            return run_ntt_pipelined(plan, mode, a, b, c, c_evals, batch, flags, st, chunk);
    }
    IoBuf ba, bb, bc, be;
    int rc;
    if ((rc = ba.init(a, bytes, true, false, st))) return rc;
    if ((rc = bb.init(mode == MODE_MUL ? b : nullptr, bcast ? plan->host.n * sizeof(u64) : bytes, true, false, st))) return rc;
    if ((rc = bc.init(c, bytes, false, true, st))) return rc;
    if ((rc = be.init(c_evals, bytes, false, true, st))) return rc;
    rc = launch_plan(plan, mode, ba.ptr<u64>(), bb.ptr<u64>(), bc.ptr<u64>(), be.ptr<u64>(), batch, flags, st);
    if (rc) return rc;
    return finish_all({&ba, &bb, &bc, &be}, st);
}
}  // namespace

namespace {
// ---- packed 32-bit wire format (q <= 2^32): halves the PCIe bytes of the host-buffer path ------------------------
__global__ void widen_u32_kernel(const u32 *__restrict__ in, u64 *__restrict__ out, size_t len4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len4; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = reinterpret_cast<const uint4 *>(in)[i];
        reinterpret_cast<ulonglong2 *>(out)[2 * i] = make_ulonglong2(v.x, v.y);
        reinterpret_cast<ulonglong2 *>(out)[2 * i + 1] = make_ulonglong2(v.z, v.w);
    }
}
__global__ void narrow_u64_kernel(const u64 *__restrict__ in, u32 *__restrict__ out, size_t len4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len4; i += (size_t)gridDim.x * blockDim.x) {
        const ulonglong2 a = reinterpret_cast<const ulonglong2 *>(in)[2 * i], b = reinterpret_cast<const ulonglong2 *>(in)[2 * i + 1];
        reinterpret_cast<uint4 *>(out)[i] = make_uint4((u32)a.x, (u32)a.y, (u32)b.x, (u32)b.y);
    }
}
inline unsigned grid_for_len(size_t work) {
    size_t g = (work + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    return (unsigned)(g < 1 ? 1 : g > cap ? cap : g);
}

// Chunked, double-buffered: [copy u32 in] -> widen -> transform -> narrow -> [copy u32 out]; host or device pointers.
int run_ntt_wire32(const fhe_ntt_plan *plan, int mode, const u32 *a, const u32 *b, u32 *c, u32 *c_evals, size_t batch,
                   int flags) {
    FHE_REQUIRE(plan != nullptr, "null plan");
    if (batch == 0) return 0;
    FHE_REQUIRE(a != nullptr && c != nullptr && (mode != MODE_MUL || b != nullptr), "null polynomial pointer");
    FHE_REQUIRE(plan->host.q <= (1ull << 32), "the 32-bit wire format needs q <= 2^32");
    FHE_REQUIRE(plan->host.n % 4 == 0, "the 32-bit wire format needs n >= 4");
    FHE_REQUIRE(!(flags & B_BROADCAST), "FHE_B_BROADCAST is not supported on the 32-bit wire");
    if (mode != MODE_MUL) b = nullptr;
    cudaStream_t st = current_stream();
    int rc = t_pipe.init();
    if (rc) return rc;
    PipeStreams &ps = t_pipe;
    const size_t n = plan->host.n;
    const size_t chunk = std::min(batch, std::max<size_t>(1, pipe_chunk_bytes() / (n * sizeof(u32))));
    const size_t w = chunk * n;  // words per chunk buffer
    const bool host = is_host_ptr(a) || (b && is_host_ptr(b)) || is_host_ptr(c) || (c_evals && is_host_ptr(c_evals));
    // per parity: u64 A, B, C, E and u32 a, b, c, e  (unused ones still reserved; at most ~0.8 GB)
    Scratch s64, s32;
    if ((rc = s64.alloc(2 * 4 * w * sizeof(u64), st))) return rc;
    if ((rc = s32.alloc(2 * 4 * w * sizeof(u32), st))) return rc;
    FHE_CUDA_OK(cudaStreamSynchronize(st));  // the scratch is used from the side streams as well
    auto b64 = [&](int which, int par) { return s64.ptr<u64>() + ((size_t)par * 4 + which) * w; };
    auto b32 = [&](int which, int par) { return s32.ptr<u32>() + ((size_t)par * 4 + which) * w; };
    size_t i = 0;
    for (size_t off = 0; off < batch && !rc; off += chunk, i++) {
        const size_t nb = std::min(chunk, batch - off), words = nb * n, bytes = words * sizeof(u32);
        const int par = (int)(i & 1);
        if (i >= 2) FHE_CUDA_OK(cudaStreamWaitEvent(ps.h2d, ps.comp_done[par], 0));  // staging of chunk i-2 consumed
        FHE_CUDA_OK(cudaMemcpyAsync(b32(0, par), a + off * n, bytes, cudaMemcpyDefault, ps.h2d));
        if (b) FHE_CUDA_OK(cudaMemcpyAsync(b32(1, par), b + off * n, bytes, cudaMemcpyDefault, ps.h2d));
        FHE_CUDA_OK(cudaEventRecord(ps.h2d_done[par], ps.h2d));
        FHE_CUDA_OK(cudaStreamWaitEvent(st, ps.h2d_done[par], 0));
        if (i >= 2) FHE_CUDA_OK(cudaStreamWaitEvent(st, ps.d2h_done[par], 0));       // outputs of chunk i-2 drained
        widen_u32_kernel<<<grid_for_len(words / 4), 256, 0, st>>>(b32(0, par), b64(0, par), words / 4);
        if (b) widen_u32_kernel<<<grid_for_len(words / 4), 256, 0, st>>>(b32(1, par), b64(1, par), words / 4);
        count_launch(b ? 2 : 1);
        if ((rc = launch_plan(plan, mode, b64(0, par), b ? b64(1, par) : nullptr, b64(2, par), c_evals ? b64(3, par) : nullptr,
                              nb, flags, st)))
            break;
        narrow_u64_kernel<<<grid_for_len(words / 4), 256, 0, st>>>(b64(2, par), b32(2, par), words / 4);
        if (c_evals) narrow_u64_kernel<<<grid_for_len(words / 4), 256, 0, st>>>(b64(3, par), b32(3, par), words / 4);
        count_launch(c_evals ? 2 : 1);
        FHE_CUDA_OK(cudaGetLastError());
        FHE_CUDA_OK(cudaEventRecord(ps.comp_done[par], st));
        FHE_CUDA_OK(cudaStreamWaitEvent(ps.d2h, ps.comp_done[par], 0));
        FHE_CUDA_OK(cudaMemcpyAsync(c + off * n, b32(2, par), bytes, cudaMemcpyDefault, ps.d2h));
        if (c_evals) FHE_CUDA_OK(cudaMemcpyAsync(c_evals + off * n, b32(3, par), bytes, cudaMemcpyDefault, ps.d2h));
        FHE_CUDA_OK(cudaEventRecord(ps.d2h_done[par], ps.d2h));
    }
    // the side streams own part of the work: the call returns with everything complete (also for device pointers)
    cudaError_t e1 = cudaStreamSynchronize(ps.d2h), e2 = cudaStreamSynchronize(ps.h2d), e3 = cudaStreamSynchronize(st);
    (void)host;
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        set_error(std::string("32-bit wire path failed: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
        return -2;
    }
    return 0;
}
}  // namespace

namespace fhe {
bool is_host_ptr(const void *p) {
    if (p == nullptr) return false;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return !(attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
}
PipeStreams &thread_pipe() { return t_pipe; }
// device-pointer transform launch for the other translation units (glwe_rq.cu)
int plan_launch(const fhe_ntt_plan *plan, int mode, const u64 *a, const u64 *b, u64 *c, u64 *c_evals, size_t batch,
                int flags, cudaStream_t st) {
    return launch_plan(plan, mode, a, b, c, c_evals, batch, flags, st);
}
}  // namespace fhe

extern "C" {

const char *fhe_last_error(void) { return t_error.c_str(); }
int fhe_device_count(int *count) {
    FHE_REQUIRE(count != nullptr, "null count");
    FHE_CUDA_OK(cudaGetDeviceCount(count));
    return 0;
}
int fhe_set_device(int device) {
    FHE_CUDA_OK(cudaSetDevice(device));
    return 0;
}
int fhe_set_stream(void *cuda_stream) {
    t_stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return 0;
}
int fhe_synchronize(void) {
    FHE_CUDA_OK(cudaStreamSynchronize(t_stream));
    return 0;
}
uint64_t fhe_launch_count(void) { return g_launches.load(); }

int fhe_ntt_plan_create(uint64_t q, uint64_t n, fhe_ntt_plan **out) {
    FHE_REQUIRE(out != nullptr, "null plan out-pointer");
    *out = nullptr;
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    device_init_once();
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto key = std::make_tuple(dev, (u64)q, (u64)n);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) {
        it->second->refs++;
        *out = it->second;
        return 0;
    }
    std::unique_ptr<fhe_ntt_plan> p(new fhe_ntt_plan());
    std::string why = build_host_tables(q, n, p->host);
    FHE_REQUIRE(why.empty(), "fhe_ntt_plan_create: " + why);
    p->device = dev;
    p->logn = hp_ilog2(n);
    p->kind = modulus_kind(q, p->logn);
    FHE_REQUIRE(p->logn <= ((p->kind == 0 || p->kind == 3) ? 15 : 14),
                "fhe_ntt_plan_create: n too large (max 2^15 for q < 2^30, 2^14 for larger q)");
    // coefficients per thread: the library default, or FHE_NTT_LOGE (a tuning knob; only values that were
    // instantiated are accepted)
    p->loge = p->kind == 0 ? LogE<Lazy32>::of(p->logn) : p->kind == 3 ? LogE<Small32>::of(p->logn)
              : p->kind == 1 ? LogE<Lazy64>::of(p->logn) : LogE<Strict64>::of(p->logn);
    if (const char *e = getenv("FHE_NTT_LOGE")) {
        const int v = atoi(e);
        const bool ok = p->kind == 0 ? ntt_loge_ok_lazy32(p->logn, v) : p->kind == 3 ? ntt_loge_ok_small32(p->logn, v)
                        : p->kind == 1 ? ntt_loge_ok_lazy64(p->logn, v) : ntt_loge_ok_strict64(p->logn, v);
        if (ok) p->loge = v;
    }
    int rc = p->kind == 0 ? upload_tables(p.get(), p->p32)
             : p->kind == 3 ? upload_tables(p.get(), p->psm)
             : p->kind == 1 ? upload_tables(p.get(), p->p64)
                            : upload_tables(p.get(), p->ps64);
    if (rc) return rc;
    p->refs = 1;
    *out = p.get();
    g_plans[key] = p.release();
    return 0;
}
void fhe_ntt_plan_destroy(fhe_ntt_plan *plan) {
    if (plan == nullptr) return;
    std::lock_guard<std::mutex> lk(g_plan_mu);
    if (--plan->refs > 0) return;
    g_plans.erase(std::make_tuple(plan->device, plan->host.q, plan->host.n));
    cudaFree(plan->d_fwd);
    cudaFree(plan->d_inv);
    delete plan;
}
int fhe_ntt_plan_info(const fhe_ntt_plan *plan, uint64_t *psi, uint64_t *n_inv, uint64_t *roots, uint64_t *roots_inv) {
    FHE_REQUIRE(plan != nullptr, "null plan");
    if (psi) *psi = plan->host.psi;
    if (n_inv) *n_inv = plan->host.n_inv;
    if (roots) memcpy(roots, plan->host.roots.data(), plan->host.n * sizeof(u64));
    if (roots_inv) memcpy(roots_inv, plan->host.roots_inv.data(), plan->host.n * sizeof(u64));
    return 0;
}
int fhe_ntt_fwd(const fhe_ntt_plan *plan, const uint64_t *in, uint64_t *out, size_t batch) {
    return run_ntt(plan, MODE_FWD, in, nullptr, out, nullptr, batch, 0);
}
int fhe_ntt_inv(const fhe_ntt_plan *plan, const uint64_t *in, uint64_t *out, size_t batch) {
    return run_ntt(plan, MODE_INV, in, nullptr, out, nullptr, batch, 0);
}
int fhe_rq_mul(const fhe_ntt_plan *plan, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t batch, int flags,
               uint64_t *c_evals) {
    return run_ntt(plan, MODE_MUL, a, b, c, c_evals, batch, flags);
}
int fhe_ntt_fwd_u32(const fhe_ntt_plan *plan, const uint32_t *in, uint32_t *out, size_t batch) {
    return run_ntt_wire32(plan, MODE_FWD, in, nullptr, out, nullptr, batch, 0);
}
int fhe_ntt_inv_u32(const fhe_ntt_plan *plan, const uint32_t *in, uint32_t *out, size_t batch) {
    return run_ntt_wire32(plan, MODE_INV, in, nullptr, out, nullptr, batch, 0);
}
int fhe_rq_mul_u32(const fhe_ntt_plan *plan, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t batch, int flags,
                   uint32_t *c_evals) {
    return run_ntt_wire32(plan, MODE_MUL, a, b, c, c_evals, batch, flags);
}

}  // extern "C"
