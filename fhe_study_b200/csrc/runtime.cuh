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
