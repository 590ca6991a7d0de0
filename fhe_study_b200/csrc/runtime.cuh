// runtime.cuh -- host-side plumbing shared by the C-ABI entry points: error state, stream selection,
// and IoBuf, which lets every entry point take either HOST or DEVICE pointers (the reference's callers
// own plain Vec<..> buffers; batch callers keep data resident in HBM).
#pragma once
#include <vector>

#include "common.cuh"

namespace fhe {

// Classifies a user pointer once; host buffers are staged through stream-ordered device scratch.
class IoBuf {
  public:
    IoBuf() {}
    ~IoBuf() { release(); }
    IoBuf(const IoBuf &) = delete;
    IoBuf &operator=(const IoBuf &) = delete;

    // in: copy host->device before use; out: copy device->host in finish().
    int init(const void *user, size_t bytes, bool in, bool out, cudaStream_t st) {
        user_ = const_cast<void *>(user);
        bytes_ = bytes;
        out_ = out;
        st_ = st;
        if (user == nullptr || bytes == 0) { dev_ = nullptr; is_host_ = false; return 0; }
        cudaPointerAttributes attr;
        cudaError_t e = cudaPointerGetAttributes(&attr, user);
        if (e != cudaSuccess) { cudaGetLastError(); attr.type = cudaMemoryTypeUnregistered; }
        if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) {
            dev_ = user_;
            is_host_ = false;
            return 0;
        }
        is_host_ = true;
        FHE_CUDA_OK(cudaMallocAsync(&dev_, bytes, st));
        owned_ = true;
        if (in) FHE_CUDA_OK(cudaMemcpyAsync(dev_, user, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    }
    template <typename T> T *ptr() const { return reinterpret_cast<T *>(dev_); }
    bool is_host() const { return is_host_; }
    // enqueue the device->host copy of an output buffer (no synchronisation)
    int finish() {
        if (is_host_ && out_ && dev_ != nullptr)
            FHE_CUDA_OK(cudaMemcpyAsync(user_, dev_, bytes_, cudaMemcpyDeviceToHost, st_));
        return 0;
    }
    void release() {
        if (owned_ && dev_ != nullptr) cudaFreeAsync(dev_, st_);
        owned_ = false;
        dev_ = nullptr;
    }

  private:
    void *user_ = nullptr, *dev_ = nullptr;
    size_t bytes_ = 0;
    bool is_host_ = false, out_ = false, owned_ = false;
    cudaStream_t st_ = nullptr;
};

// Stream-ordered device scratch with scope lifetime (freed on every exit path).
class Scratch {
  public:
    Scratch() {}
    ~Scratch() { if (p_ != nullptr) cudaFreeAsync(p_, st_); }
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
    int alloc(size_t bytes, cudaStream_t st) {
        st_ = st;
        FHE_CUDA_OK(cudaMallocAsync(&p_, bytes ? bytes : 1, st));
        return 0;
    }
    template <typename T> T *ptr() const { return reinterpret_cast<T *>(p_); }

  private:
    void *p_ = nullptr;
    cudaStream_t st_ = nullptr;
};

// Host-buffer path for large batches: the batch is cut into chunks and H2D copies, kernels and D2H copies of
// consecutive chunks overlap on three streams (PCIe is full duplex), double-buffered on the device.
struct PipeStreams {
    cudaStream_t h2d = nullptr, d2h = nullptr, comp2 = nullptr;  // comp2: second compute stream (odd chunks)
    cudaEvent_t h2d_done[2] = {nullptr, nullptr}, comp_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
    int device = -1;
    int init() {
        int dev = 0;
        FHE_CUDA_OK(cudaGetDevice(&dev));
        if (device == dev) return 0;
        if (device >= 0) {  // the thread moved to another device: the old streams and events belong to the old one
            int cur = dev;
            if (cudaSetDevice(device) == cudaSuccess) {
                cudaStreamDestroy(h2d); cudaStreamDestroy(d2h); cudaStreamDestroy(comp2);
                for (int i = 0; i < 2; i++) { cudaEventDestroy(h2d_done[i]); cudaEventDestroy(comp_done[i]); cudaEventDestroy(d2h_done[i]); }
            }
            cudaSetDevice(cur);
            cudaGetLastError();
            device = -1;
        }
        FHE_CUDA_OK(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
        FHE_CUDA_OK(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
        FHE_CUDA_OK(cudaStreamCreateWithFlags(&comp2, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            FHE_CUDA_OK(cudaEventCreateWithFlags(&h2d_done[i], cudaEventDisableTiming));
            FHE_CUDA_OK(cudaEventCreateWithFlags(&comp_done[i], cudaEventDisableTiming));
            FHE_CUDA_OK(cudaEventCreateWithFlags(&d2h_done[i], cudaEventDisableTiming));
        }
        device = dev;
        return 0;
    }
};
PipeStreams &thread_pipe();  // per host thread (lib_core.cu)

// Host-buffer batches: `batch` independent units cut into chunks; H2D copy of chunk i+1, fn(dev_in, dev_out, nb, stream, parity)
// of chunk i and D2H copy of chunk i-1 overlap (double-buffered device staging).  Even and odd chunks compute on two
// different streams, so the tail of one chunk's kernels (a partial last wave) overlaps the head of the next chunk's.
// fn must keep any scratch it uses per parity.  Returns with `out` complete.
template <class F>
int run_host_pipelined(const void *in, size_t in_unit, void *out, size_t out_unit, size_t batch, size_t chunk, cudaStream_t st,
                       F fn) {
    PipeStreams &ps = thread_pipe();
    int rc = ps.init();
    if (rc) return rc;
    Scratch sin, sout;
    if ((rc = sin.alloc(2 * chunk * in_unit, st))) return rc;
    if ((rc = sout.alloc(2 * chunk * out_unit, st))) return rc;
    FHE_CUDA_OK(cudaStreamSynchronize(st));  // the scratch (and the caller's) is used from the side streams as well
    // a failing CUDA call inside the loop must not return: the side streams may still have work queued on the staging
    // buffers (freed by the Scratch destructors), so the loop is left and the common tail synchronises every stream first
#define FHE_PIPE_TRY(expr)                                                              \
    {                                                                                   \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));              \
            rc = -2;                                                                    \
            break;                                                                      \
        }                                                                               \
    }
    size_t i = 0;
    for (size_t off = 0; off < batch && !rc; off += chunk, i++) {
        const size_t nb = batch - off < chunk ? batch - off : chunk;
        const int par = (int)(i & 1);
        cudaStream_t cs = par ? ps.comp2 : st;
        char *din = sin.ptr<char>() + (size_t)par * chunk * in_unit, *dout = sout.ptr<char>() + (size_t)par * chunk * out_unit;
        if (i >= 2) FHE_PIPE_TRY(cudaStreamWaitEvent(ps.h2d, ps.comp_done[par], 0));  // staging of chunk i-2 consumed
        FHE_PIPE_TRY(cudaMemcpyAsync(din, (const char *)in + off * in_unit, nb * in_unit, cudaMemcpyHostToDevice, ps.h2d));
        FHE_PIPE_TRY(cudaEventRecord(ps.h2d_done[par], ps.h2d));
        FHE_PIPE_TRY(cudaStreamWaitEvent(cs, ps.h2d_done[par], 0));
        if (i >= 2) FHE_PIPE_TRY(cudaStreamWaitEvent(cs, ps.d2h_done[par], 0));       // outputs of chunk i-2 drained
        if ((rc = fn(din, dout, nb, cs, par))) break;
        FHE_PIPE_TRY(cudaEventRecord(ps.comp_done[par], cs));
        FHE_PIPE_TRY(cudaStreamWaitEvent(ps.d2h, ps.comp_done[par], 0));
        FHE_PIPE_TRY(cudaMemcpyAsync((char *)out + off * out_unit, dout, nb * out_unit, cudaMemcpyDeviceToHost, ps.d2h));
        FHE_PIPE_TRY(cudaEventRecord(ps.d2h_done[par], ps.d2h));
    }
#undef FHE_PIPE_TRY
    cudaError_t e1 = cudaStreamSynchronize(ps.d2h), e2 = cudaStreamSynchronize(ps.h2d), e3 = cudaStreamSynchronize(st),
                e4 = cudaStreamSynchronize(ps.comp2);
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess) {
        set_error(std::string("pipelined transfer failed: ") +
                  cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3 != cudaSuccess ? e3 : e4));
        return -2;
    }
    return 0;
}
bool is_host_ptr(const void *p);

// Scope helper: finish() every buffer, and synchronise the stream iff any of them lives on the host
// (the call then has the reference's blocking semantics; all-device calls stay asynchronous).
inline int finish_all(std::initializer_list<IoBuf *> bufs, cudaStream_t st) {
    bool any_host = false;
    for (IoBuf *b : bufs) {
        int rc = b->finish();
        if (rc) return rc;
        any_host |= b->is_host();
    }
    if (any_host) FHE_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

}  // namespace fhe
