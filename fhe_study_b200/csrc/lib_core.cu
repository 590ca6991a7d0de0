// lib_core.cu -- runtime state (errors, stream, launch counter) and the NTT-plan / Rq entry points of the
// C ABI declared in include/fhe_b200.h.
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <memory>
#include <tuple>
#include <type_traits>

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
    static std::atomic<unsigned long long> done_mask{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    if ((done_mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    cudaGetLastError();
    done_mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
}
int num_sms() {
    static std::atomic<int> sms[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int v = sms[dev & 63].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
        sms[dev & 63].store(v, std::memory_order_relaxed);
    }
    return v;
}
// A handle (plan, key, ...) belongs to the device it was created on: using it while another device is current would
// launch kernels that read the other GPU's memory (an illegal-address fault without peer access).
int check_device(int handle_device, const char *what) {
    int dev = -1;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    if (dev != handle_device) {
        set_error(std::string(what) + ": handle was created on device " + std::to_string(handle_device) +
                  " but device " + std::to_string(dev) + " is current (fhe_set_device)");
        return -1;
    }
    return 0;
}

}  // namespace fhe

using namespace fhe;

namespace {
std::mutex g_plan_mu;
std::map<std::tuple<int, u64, u64>, fhe_ntt_plan *> g_plans;

template <class M> int upload_tables(fhe_ntt_plan *p, NttParams<M> &dst, void **d_fwd = nullptr, void **d_inv = nullptr) {
    if (d_fwd == nullptr) d_fwd = &p->d_fwd;
    if (d_inv == nullptr) d_inv = &p->d_inv;
    ExpandedTables<M> x;
    expand_tables(p->host, x, p->loge);
    const size_t bytes = sizeof(typename M::T) * p->host.n;
    FHE_CUDA_OK(cudaMalloc(d_fwd, bytes));
    FHE_CUDA_OK(cudaMalloc(d_inv, bytes));
    FHE_CUDA_OK(cudaMemcpy(*d_fwd, x.fwd.data(), bytes, cudaMemcpyHostToDevice));
    FHE_CUDA_OK(cudaMemcpy(*d_inv, x.inv.data(), bytes, cudaMemcpyHostToDevice));
    dst.mod = x.mod;
    dst.fwd = reinterpret_cast<const typename M::T *>(*d_fwd);
    dst.inv = reinterpret_cast<const typename M::T *>(*d_inv);
    dst.ninv = x.ninv;
    dst.s_ninv = x.s_ninv;
    dst.ninv_pw = x.ninv_pw;
    dst.s_ninv_pw = x.s_ninv_pw;
    for (u64 i = 0; i < 64; i++) {
        dst.c_fwd[i] = x.fwd[i < p->host.n ? i : 0];
        dst.c_inv[i] = x.inv[i < p->host.n ? i : 0];
    }
    dst.fwdw = dst.invw = nullptr;
    if (!x.fwdw.empty()) {  // radix-4 policies: the 4-byte twiddle tables
        const size_t wb = sizeof(u32) * p->host.n;
        FHE_CUDA_OK(cudaMalloc(&p->d_fwd4w, wb));
        FHE_CUDA_OK(cudaMalloc(&p->d_inv4w, wb));
        FHE_CUDA_OK(cudaMemcpy(p->d_fwd4w, x.fwdw.data(), wb, cudaMemcpyHostToDevice));
        FHE_CUDA_OK(cudaMemcpy(p->d_inv4w, x.invw.data(), wb, cudaMemcpyHostToDevice));
        dst.fwdw = reinterpret_cast<const u32 *>(p->d_fwd4w);
        dst.invw = reinterpret_cast<const u32 *>(p->d_inv4w);
    }
    return 0;
}

// polymul of two coefficient-form operands: the dual-operand kernel where the plan says so (FHE_NTT_DUAL overrides)
inline int mul_mode_for(const fhe_ntt_plan *plan, int mode, int flags) {
    if (mode != MODE_MUL || (flags & (A_IS_EVALS | B_IS_EVALS))) return mode;
    return plan->gpark ? MODE_MULG : plan->staged ? MODE_MULS : plan->dual ? MODE_MUL2 : mode;
}
int launch_plan(const fhe_ntt_plan *plan, int mode, const u64 *a, const u64 *b, u64 *c, u64 *c_evals, size_t batch,
                int flags, cudaStream_t st) {
    int rc;
    mode = mul_mode_for(plan, mode, flags);
    if (mode == MODE_MULS && plan->staged == 2) flags |= STAGE_TMA;
    if (plan->fermat) {  // q = 65537: radix-4 butterflies (kind 3, and kind 0 at n = 2^15)
        rc = ntt_launch_fermat32(plan->logn, plan->loge, mode, plan->pfm, a, b, c, c_evals, batch, flags, st);
        if (!rc) count_launch(1);
        return rc;
    }
    switch (plan->kind) {
        case 3: rc = ntt_launch_small32(plan->logn, plan->loge, mode, plan->psm, a, b, c, c_evals, batch, flags, st); break;
        case 0: rc = ntt_launch_lazy32(plan->logn, plan->loge, mode, plan->p32, a, b, c, c_evals, batch, flags, st); break;
        case 1: rc = ntt_launch_lazy64(plan->logn, plan->loge, mode, plan->p64, a, b, c, c_evals, batch, flags, st); break;
        default: rc = ntt_launch_strict64(plan->logn, plan->loge, mode, plan->ps64, a, b, c, c_evals, batch, flags, st);
    }
    if (!rc) count_launch(1);
    return rc;
}
// packed 32-bit words in global memory (q <= 2^32): the same kernels instantiated with u32 loads and stores
int launch_plan(const fhe_ntt_plan *plan, int mode, const u32 *a, const u32 *b, u32 *c, u32 *c_evals, size_t batch,
                int flags, cudaStream_t st) {
    int rc;
    mode = mul_mode_for(plan, mode, flags);
    if (mode == MODE_MULS && plan->staged == 2) flags |= STAGE_TMA;
    if (plan->fermat) {  // q = 65537: radix-4 butterflies (kind 3, and kind 0 at n = 2^15)
        rc = ntt_launch_fermat32_u32(plan->logn, plan->loge, mode, plan->pfm, a, b, c, c_evals, batch, flags, st);
        if (!rc) count_launch(1);
        return rc;
    }
    switch (plan->kind) {
        case 3: rc = ntt_launch_small32_u32(plan->logn, plan->loge, mode, plan->psm, a, b, c, c_evals, batch, flags, st); break;
        case 0: rc = ntt_launch_lazy32_u32(plan->logn, plan->loge, mode, plan->p32, a, b, c, c_evals, batch, flags, st); break;
        case 1: rc = ntt_launch_lazy64_u32(plan->logn, plan->loge, mode, plan->p64, a, b, c, c_evals, batch, flags, st); break;
        default: set_error("the 32-bit word format needs q <= 2^32"); return -1;
    }
    if (!rc) count_launch(1);
    return rc;
}

// bit-packed words (q < 2^30, n >= 1024): `bits` rides in the upper bits of the kernel's flags word
constexpr int PACKED_BITS_SHIFT = 8;
int launch_plan(const fhe_ntt_plan *plan, int mode, const pk32 *a, const pk32 *b, pk32 *c, pk32 *c_evals, size_t batch,
                int flags, cudaStream_t st) {
    int rc;
    mode = mul_mode_for(plan, mode, flags);
    if (mode == MODE_MULS && plan->staged == 2) flags |= STAGE_TMA;
    if (plan->fermat) {  // q = 65537: radix-4 butterflies (kind 3, and kind 0 at n = 2^15)
        rc = ntt_launch_fermat32_pk(plan->logn, plan->loge, mode, plan->pfm, a, b, c, c_evals, batch, flags, st);
        if (!rc) count_launch(1);
        return rc;
    }
    switch (plan->kind) {
        case 3: rc = ntt_launch_small32_pk(plan->logn, plan->loge, mode, plan->psm, a, b, c, c_evals, batch, flags, st); break;
        case 0: rc = ntt_launch_lazy32_pk(plan->logn, plan->loge, mode, plan->p32, a, b, c, c_evals, batch, flags, st); break;
        default: set_error("the bit-packed format needs q < 2^30"); return -1;
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

// inside the chunk loops a failing CUDA call must not return: the side streams may still have work queued on the
// staging buffers, so the loop is left and the common tail synchronises all streams before anything is freed
#define FHE_PIPE_OK(expr)                                                               \
    {                                                                                   \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
            rc = -2;                                                                    \
            break;                                                                      \
        }                                                                               \
    }

// Host-buffer batches of NTT / INTT / polymul, W = u64, u32 or bit-packed words (`row` words per polynomial):
// chunked, double-buffered on the device; the H2D copy of chunk i+1, the ONE kernel launch of chunk i and the D2H copy
// of chunk i-1 overlap on three streams.
template <typename W>
int run_ntt_pipelined(const fhe_ntt_plan *plan, int mode, const W *a, const W *b, W *c, W *c_evals, size_t batch,
                      int flags, cudaStream_t st, size_t chunk, size_t row) {
    int rc = t_pipe.init();
    if (rc) return rc;
    PipeStreams &ps = t_pipe;
    const size_t cbytes = chunk * row * sizeof(W);
    const int nbuf = 1 + (b ? 1 : 0) + 1 + (c_evals ? 1 : 0);
    Scratch scratch;
    if ((rc = scratch.alloc(2 * nbuf * cbytes, st))) return rc;
    W *dev = scratch.ptr<W>();
    FHE_CUDA_OK(cudaStreamSynchronize(st));  // the scratch is used from the side streams as well
    auto buf = [&](int which, int parity) { return dev + ((size_t)parity * nbuf + which) * chunk * row; };
    const int ib = 1, ic = b ? 2 : 1, ie = ic + 1;
    size_t i = 0;
    for (size_t off = 0; off < batch; off += chunk, i++) {
        const size_t nb = std::min(chunk, batch - off), bytes = nb * row * sizeof(W);
        const int par = (int)(i & 1);
        if (i >= 2) FHE_PIPE_OK(cudaStreamWaitEvent(ps.h2d, ps.comp_done[par], 0));  // inputs of chunk i-2 consumed
        FHE_PIPE_OK(cudaMemcpyAsync(buf(0, par), a + off * row, bytes, cudaMemcpyHostToDevice, ps.h2d));
        if (b) FHE_PIPE_OK(cudaMemcpyAsync(buf(ib, par), b + off * row, bytes, cudaMemcpyHostToDevice, ps.h2d));
        FHE_PIPE_OK(cudaEventRecord(ps.h2d_done[par], ps.h2d));
        FHE_PIPE_OK(cudaStreamWaitEvent(st, ps.h2d_done[par], 0));
        if (i >= 2) FHE_PIPE_OK(cudaStreamWaitEvent(st, ps.d2h_done[par], 0));       // outputs of chunk i-2 drained
        if ((rc = launch_plan(plan, mode, buf(0, par), b ? buf(ib, par) : (const W *)nullptr, buf(ic, par),
                              c_evals ? buf(ie, par) : (W *)nullptr, nb, flags, st)))
            break;
        FHE_PIPE_OK(cudaEventRecord(ps.comp_done[par], st));
        FHE_PIPE_OK(cudaStreamWaitEvent(ps.d2h, ps.comp_done[par], 0));
        FHE_PIPE_OK(cudaMemcpyAsync(c + off * row, buf(ic, par), bytes, cudaMemcpyDeviceToHost, ps.d2h));
        if (c_evals) FHE_PIPE_OK(cudaMemcpyAsync(c_evals + off * row, buf(ie, par), bytes, cudaMemcpyDeviceToHost, ps.d2h));
        FHE_PIPE_OK(cudaEventRecord(ps.d2h_done[par], ps.d2h));
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

// NTT / INTT / polymul over words of type W (u64: SURVEY 8b layout; u32: packed words for q <= 2^32; pk32: `bits`
// bits per coefficient).  Every pointer may be a host or a device pointer; all-device calls are ONE asynchronous
// kernel launch on the current stream.
template <typename W>
int run_ntt(const fhe_ntt_plan *plan, int mode, const W *a, const W *b, W *c, W *c_evals, size_t batch, int flags,
            int bits = 0) {
    FHE_REQUIRE(plan != nullptr, "null plan");
    if (batch == 0) return 0;
    FHE_REQUIRE(a != nullptr && c != nullptr && (mode != MODE_MUL || b != nullptr), "null polynomial pointer");
    FHE_REQUIRE((flags & ~7) == 0, "unknown flag bits");
    size_t row = plan->host.n;  // words per polynomial
    if (std::is_same<W, u32>::value) FHE_REQUIRE(plan->host.q <= (1ull << 32), "the 32-bit word format needs q <= 2^32");
    if (std::is_same<W, pk32>::value) {
        FHE_REQUIRE(bits >= 16 && bits <= 31, "bit-packed format: 16 <= bits <= 31");
        FHE_REQUIRE(plan->host.q <= (1ull << bits), "bit-packed format: q does not fit the field width");
        FHE_REQUIRE(plan->host.n >= 1024, "the bit-packed format needs n >= 1024");
        row = plan->host.n / 32 * (size_t)bits;
        flags |= bits << PACKED_BITS_SHIFT;
    }
    int rc = check_device(plan->device, "NTT plan");
    if (rc) return rc;
    cudaStream_t st = current_stream();
    const size_t bytes = batch * row * sizeof(W);
    if (mode != MODE_MUL) b = nullptr;
    const bool bcast = mode == MODE_MUL && (flags & B_BROADCAST);
    if (!bcast) {   // all-host call on a batch worth pipelining (>= 4 chunks of ~32 MiB per operand)
        const size_t chunk = std::max<size_t>(1, pipe_chunk_bytes() / (row * sizeof(W)));
        if (batch >= 4 * chunk && is_host_ptr(a) && (!b || is_host_ptr(b)) && is_host_ptr(c) &&
            (!c_evals || is_host_ptr(c_evals)) && a != c && b != c)
            return run_ntt_pipelined<W>(plan, mode, a, b, c, c_evals, batch, flags, st, chunk, row);
    }
    IoBuf ba, bb, bc, be;
    if ((rc = ba.init(a, bytes, true, false, st))) return rc;
    if ((rc = bb.init(mode == MODE_MUL ? b : nullptr, bcast ? row * sizeof(W) : bytes, true, false, st))) return rc;
    if ((rc = bc.init(c, bytes, false, true, st))) return rc;
    if ((rc = be.init(c_evals, bytes, false, true, st))) return rc;
    rc = launch_plan(plan, mode, ba.ptr<W>(), bb.ptr<W>(), bc.ptr<W>(), be.ptr<W>(), batch, flags, st);
    if (rc) return rc;
    return finish_all({&ba, &bb, &bc, &be}, st);
}
}  // namespace

namespace {
__global__ void first_noncanonical_kernel(const u64 *__restrict__ w, size_t len, u64 q, unsigned long long *first) {
    unsigned long long best = ~0ull;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x)
        if (w[i] >= q && i < best) best = i;
    if (best != ~0ull) atomicMin(first, best);
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
    int rc = check_device(plan->device, "NTT plan");
    if (rc) return rc;
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
int fhe_current_device(int *device) {
    FHE_REQUIRE(device != nullptr, "null device out-pointer");
    FHE_CUDA_OK(cudaGetDevice(device));
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
    // Which polymul kernel serves two coefficient-form operands (measured on B200, q = 65537, M polymul/s):
    //   dual-operand (MODE_MUL2): N=2048 94.9 -> 99.7; slower elsewhere (N=1024 217 -> 199, N=16384 6.75 -> 5.84)
    //   NTT(a) parked in the output row (MODE_MULG): N=16384 6.75 -> 7.76 (two 64-register CTAs per SM); N=8192 unchanged
    //   persistent + cp.async operand staging (MODE_MULS): N=16384 6.75 -> 6.93, N=8192 18.7 -> 15.5: off
    //   (FHE_NTT_STAGED=2: the same loop with the operands posted as TMA bulk copies by one thread)
    // FHE_NTT_DUAL / FHE_NTT_GPARK / FHE_NTT_STAGED = 0|1 override (tuning knobs and tests).
    const bool w32 = p->kind == 0 || p->kind == 3;
    p->dual = w32 && p->logn == 11;
    if (const char *e = getenv("FHE_NTT_DUAL")) p->dual = atoi(e) != 0;
    p->gpark = (w32 && p->logn >= 13) || (!w32 && p->logn == 13);  // 32-bit N=8192: unchanged for the radix-2 kernels, 19.5 -> 19.9 M/s under Fermat32  // 62-bit q, N=8192: 3.89 -> 4.05 M/s; N=16384: 1.86 -> 1.83 (off)
    if (const char *e = getenv("FHE_NTT_GPARK")) p->gpark = atoi(e) != 0;
    p->staged = 0;
    if (const char *e = getenv("FHE_NTT_STAGED")) p->staged = atoi(e) < 0 ? 0 : atoi(e) > 2 ? 2 : atoi(e);  // 1: cp.async (LDGSTS), 2: bulk copies (TMA)
    int rc = p->kind == 0 ? upload_tables(p.get(), p->p32)
             : p->kind == 3 ? upload_tables(p.get(), p->psm)
             : p->kind == 1 ? upload_tables(p.get(), p->p64)
                            : upload_tables(p.get(), p->ps64);
    if (rc) return rc;
    // q = 65537 (the reference's modulus): radix-4 butterflies, three twiddle products and a shift per four butterflies
    // (Fermat32).  FHE_NTT_FERMAT=0 keeps the radix-2 Small32 kernels (A/B measurements and tests).
    p->fermat = (p->kind == 3 || p->kind == 0) && fermat_ok(p->host);  // kind 0: n = 2^15, where 2q * n > 2^32 rules Small32 out
    if (const char *e = getenv("FHE_NTT_FERMAT")) p->fermat = p->fermat && atoi(e) != 0;
    if (p->fermat && (rc = upload_tables(p.get(), p->pfm, &p->d_fwd4, &p->d_inv4))) return rc;
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
    cudaFree(plan->d_fwd4);
    cudaFree(plan->d_inv4);
    cudaFree(plan->d_fwd4w);
    cudaFree(plan->d_inv4w);
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
int fhe_rq_check_canonical(uint64_t q, const uint64_t *words, size_t len, uint64_t *first_bad) {
    FHE_REQUIRE(first_bad != nullptr && (words != nullptr || len == 0), "fhe_rq_check_canonical: null pointer");
    *first_bad = ~0ull;
    if (len == 0) return 0;
    cudaStream_t st = current_stream();
    IoBuf bw;
    int rc = bw.init(words, len * sizeof(u64), true, false, st);
    if (rc) return rc;
    Scratch res;
    if ((rc = res.alloc(sizeof(unsigned long long), st))) return rc;
    FHE_CUDA_OK(cudaMemsetAsync(res.ptr<void>(), 0xff, sizeof(unsigned long long), st));
    const size_t g = (len + 255) / 256, cap = (size_t)num_sms() * 16;
    first_noncanonical_kernel<<<(unsigned)(g > cap ? cap : g), 256, 0, st>>>(bw.ptr<u64>(), len, q, res.ptr<unsigned long long>());
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    FHE_CUDA_OK(cudaMemcpyAsync(first_bad, res.ptr<void>(), sizeof(u64), cudaMemcpyDeviceToHost, st));
    FHE_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}
int fhe_ntt_plan_config(const fhe_ntt_plan *plan, int *config) {
    FHE_REQUIRE(plan != nullptr && config != nullptr, "null plan or config");
    config[0] = plan->fermat ? 4 : plan->kind;
    config[1] = plan->loge;
    config[2] = plan->dual;
    config[3] = plan->gpark;
    config[4] = plan->staged;
    return 0;
}
int fhe_ntt_fwd(const fhe_ntt_plan *plan, const uint64_t *in, uint64_t *out, size_t batch) {
    return run_ntt<u64>(plan, MODE_FWD, in, nullptr, out, nullptr, batch, 0);
}
int fhe_ntt_inv(const fhe_ntt_plan *plan, const uint64_t *in, uint64_t *out, size_t batch) {
    return run_ntt<u64>(plan, MODE_INV, in, nullptr, out, nullptr, batch, 0);
}
int fhe_rq_mul(const fhe_ntt_plan *plan, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t batch, int flags,
               uint64_t *c_evals) {
    return run_ntt<u64>(plan, MODE_MUL, a, b, c, c_evals, batch, flags);
}
int fhe_ntt_fwd_packed(const fhe_ntt_plan *plan, int bits, const uint32_t *in, uint32_t *out, size_t batch) {
    return run_ntt<pk32>(plan, MODE_FWD, reinterpret_cast<const pk32 *>(in), nullptr, reinterpret_cast<pk32 *>(out), nullptr, batch, 0, bits);
}
int fhe_ntt_inv_packed(const fhe_ntt_plan *plan, int bits, const uint32_t *in, uint32_t *out, size_t batch) {
    return run_ntt<pk32>(plan, MODE_INV, reinterpret_cast<const pk32 *>(in), nullptr, reinterpret_cast<pk32 *>(out), nullptr, batch, 0, bits);
}
int fhe_rq_mul_packed(const fhe_ntt_plan *plan, int bits, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t batch,
                      int flags, uint32_t *c_evals) {
    return run_ntt<pk32>(plan, MODE_MUL, reinterpret_cast<const pk32 *>(a), reinterpret_cast<const pk32 *>(b),
                         reinterpret_cast<pk32 *>(c), reinterpret_cast<pk32 *>(c_evals), batch, flags, bits);
}
// host-side (de)serialisation of the bit-packed format: `len` coefficients <-> len * bits / 32 words (len % 32 == 0)
int fhe_pack_bits(int bits, const uint64_t *in, uint32_t *out, size_t len) {
    FHE_REQUIRE(bits >= 1 && bits <= 32 && in != nullptr && out != nullptr && len % 32 == 0, "fhe_pack_bits: bad arguments");
    const size_t blocks = len / 32;
    for (size_t blk = 0; blk < blocks; blk++) {
        const uint64_t *src = in + blk * 32;
        uint32_t *dst = out + blk * (size_t)bits;
        uint64_t acc = 0;
        int have = 0;
        size_t w = 0;
        for (int i = 0; i < 32; i++) {
            FHE_REQUIRE(bits == 32 ? src[i] <= 0xffffffffull : (src[i] >> bits) == 0, "fhe_pack_bits: coefficient does not fit the field");
            acc |= src[i] << have;
            have += bits;
            if (have >= 32) {
                dst[w++] = (uint32_t)acc;
                acc >>= 32;
                have -= 32;
            }
        }
    }
    return 0;
}
int fhe_unpack_bits(int bits, const uint32_t *in, uint64_t *out, size_t len) {
    FHE_REQUIRE(bits >= 1 && bits <= 32 && in != nullptr && out != nullptr && len % 32 == 0, "fhe_unpack_bits: bad arguments");
    const uint64_t mask = bits == 32 ? 0xffffffffull : ((1ull << bits) - 1);
    const size_t blocks = len / 32;
    for (size_t blk = 0; blk < blocks; blk++) {
        const uint32_t *src = in + blk * (size_t)bits;
        uint64_t *dst = out + blk * 32;
        uint64_t acc = 0;
        int have = 0;
        size_t w = 0;
        for (int i = 0; i < 32; i++) {
            if (have < bits) {
                acc |= (uint64_t)src[w++] << have;
                have += 32;
            }
            dst[i] = acc & mask;
            acc >>= bits;
            have -= bits;
        }
    }
    return 0;
}
int fhe_ntt_fwd_u32(const fhe_ntt_plan *plan, const uint32_t *in, uint32_t *out, size_t batch) {
    return run_ntt<u32>(plan, MODE_FWD, in, nullptr, out, nullptr, batch, 0);
}
int fhe_ntt_inv_u32(const fhe_ntt_plan *plan, const uint32_t *in, uint32_t *out, size_t batch) {
    return run_ntt<u32>(plan, MODE_INV, in, nullptr, out, nullptr, batch, 0);
}
int fhe_rq_mul_u32(const fhe_ntt_plan *plan, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t batch, int flags,
                   uint32_t *c_evals) {
    return run_ntt<u32>(plan, MODE_MUL, a, b, c, c_evals, batch, flags);
}

}  // extern "C"
