// lib_tlwe.cu -- C-ABI entry points of the TLWE path: key switch, sample extraction, mod switch, blind rotation
// and bootstrapping (tfhe/src/tlwe.rs:101-161, tfhe/src/tglwe.rs:89-119).
#include <algorithm>
#include <memory>
#include <vector>

#include "../../include/fhe_b200.h"
#include "runtime.cuh"
#include "tlwe.cuh"
#include "torus.cuh"

using namespace fhe;

struct fhe_ksk {
    Ksk k;
};

extern "C" {

// shared tail of fhe_ksk_load / fhe_ksk_generate: the handle owns `rows` (already on the device)
static int ksk_finish_handle(std::unique_ptr<fhe_ksk> &h, cudaStream_t st, fhe_ksk **out) {
    if (h->k.l == 64 && h->k.kn_in % 2 == 0 && h->k.kn_in <= 131072) {
        int rc = ksk_build_mma_layout(h->k, st);
        if (rc) {
            cudaFree(h->k.rows);
            cudaFree(h->k.mma_blocks);
            cudaFree(h->k.bcol);
            return rc;
        }
    }
    *out = h.release();
    return 0;
}

int fhe_ksk_load(uint64_t kn_in, uint64_t kn_out, uint64_t l, const uint64_t *rows, fhe_ksk **out) {
    FHE_REQUIRE(out && rows, "null pointer");
    *out = nullptr;
    FHE_REQUIRE(kn_in >= 1 && kn_out >= 1 && l >= 1 && l <= 64, "fhe_ksk_load: need kn_in, kn_out >= 1 and 1 <= l <= 64");
    device_init_once();
    std::unique_ptr<fhe_ksk> h(new fhe_ksk());
    h->k.kn_in = kn_in;
    h->k.kn_out = kn_out;
    h->k.l = l;
    const size_t bytes = kn_in * l * (kn_out + 1) * sizeof(u64);
    FHE_CUDA_OK(cudaMalloc((void **)&h->k.rows, bytes));
    cudaStream_t st = current_stream();
    // host or device source; the handle owns its own resident copy
    cudaError_t e = cudaMemcpyAsync(h->k.rows, rows, bytes, cudaMemcpyDefault, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(h->k.rows);
        set_error(std::string("fhe_ksk_load: ") + cudaGetErrorString(e));
        return -2;
    }
    return ksk_finish_handle(h, st, out);
}

// KSK generated on the device (tlwe.rs:84-100): no 537 MB upload, no CPU sampling loop.
int fhe_ksk_generate(uint64_t kn_in, uint64_t kn_out, uint64_t l, const uint64_t *sk, const uint64_t *new_sk, double sigma,
                     uint64_t seed, int uniform_mask, fhe_ksk **out) {
    FHE_REQUIRE(out && sk && new_sk, "null pointer");
    *out = nullptr;
    FHE_REQUIRE(kn_in >= 1 && kn_out >= 1 && l >= 1 && l <= 64, "fhe_ksk_generate: need kn_in, kn_out >= 1 and 1 <= l <= 64");
    device_init_once();
    cudaStream_t st = current_stream();
    IoBuf bs, bn;
    int rc;
    if ((rc = bs.init(sk, kn_in * 8, true, false, st))) return rc;
    if ((rc = bn.init(new_sk, kn_out * 8, true, false, st))) return rc;
    std::unique_ptr<fhe_ksk> h(new fhe_ksk());
    h->k.kn_in = kn_in;
    h->k.kn_out = kn_out;
    h->k.l = l;
    FHE_CUDA_OK(cudaMalloc((void **)&h->k.rows, kn_in * l * (kn_out + 1) * sizeof(u64)));
    rc = ksk_generate_device(h->k.rows, bs.ptr<u64>(), bn.ptr<u64>(), seed, (u32)kn_in, (u32)kn_out, (u32)l, sigma,
                             uniform_mask != 0, st);
    cudaError_t e = rc ? cudaSuccess : cudaStreamSynchronize(st);
    if (rc || e != cudaSuccess) {
        cudaFree(h->k.rows);
        if (!rc) set_error(std::string("fhe_ksk_generate: ") + cudaGetErrorString(e));
        return rc ? rc : -2;
    }
    return ksk_finish_handle(h, st, out);
}
// copies the key rows (kn_in * l * (kn_out + 1) words, the layout fhe_ksk_load takes) back out of a handle
int fhe_ksk_export(const fhe_ksk *h, uint64_t *rows) {
    FHE_REQUIRE(h != nullptr && rows != nullptr, "null pointer");
    cudaStream_t st = current_stream();
    const size_t bytes = h->k.kn_in * h->k.l * (h->k.kn_out + 1) * sizeof(u64);
    FHE_CUDA_OK(cudaMemcpyAsync(rows, h->k.rows, bytes, cudaMemcpyDefault, st));
    FHE_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}
void fhe_ksk_destroy(fhe_ksk *h) {
    if (!h) return;
    cudaFree(h->k.rows);
    cudaFree(h->k.mma_blocks);
    cudaFree(h->k.bcol);
    delete h;
}

int fhe_key_switch(const fhe_ksk *h, const uint64_t *ct, uint64_t *out, size_t batch) {
    FHE_REQUIRE(h != nullptr, "null KSK handle");
    if (batch == 0) return 0;
    FHE_REQUIRE(ct && out, "null ciphertext pointer");
    cudaStream_t st = current_stream();
    const size_t chunk = std::max<size_t>(256, ((16u << 20) / ((h->k.kn_in + 1) * 8)) / 256 * 256);  // ~16 MB stages
    if (batch >= 4 * chunk && is_host_ptr(ct) && is_host_ptr(out))  // all-host call: copies and kernels overlap
        return run_host_pipelined(ct, (h->k.kn_in + 1) * 8, out, (h->k.kn_out + 1) * 8, batch, chunk, st,
                                  [&](const void *din, void *dout, size_t nb, cudaStream_t s, int) {
                                      return key_switch_device(h->k, (const u64 *)din, (u64 *)dout, nb, s);
                                  });
    IoBuf bi, bo;
    int rc;
    if ((rc = bi.init(ct, batch * (h->k.kn_in + 1) * 8, true, false, st))) return rc;
    if ((rc = bo.init(out, batch * (h->k.kn_out + 1) * 8, false, true, st))) return rc;
    if ((rc = key_switch_device(h->k, bi.ptr<u64>(), bo.ptr<u64>(), batch, st))) return rc;
    return finish_all({&bi, &bo}, st);
}

// TLWE::encrypt_s for `batch` already-encoded messages (counter-based sampler, row b of the stream = ciphertext b)
int fhe_tlwe_encrypt(uint64_t kn, const uint64_t *sk, const uint64_t *msgs, double sigma, uint64_t seed, int uniform_mask,
                     uint64_t *ct, size_t batch) {
    if (batch == 0) return 0;
    FHE_REQUIRE(sk && msgs && ct, "fhe_tlwe_encrypt: null pointer");
    FHE_REQUIRE(kn >= 1 && kn < (1ull << 31), "fhe_tlwe_encrypt: kn out of range");
    cudaStream_t st = current_stream();
    IoBuf bs, bm, bo;
    int rc;
    if ((rc = bs.init(sk, kn * 8, true, false, st))) return rc;
    if ((rc = bm.init(msgs, batch * 8, true, false, st))) return rc;
    if ((rc = bo.init(ct, batch * (kn + 1) * 8, false, true, st))) return rc;
    if ((rc = tlwe_encrypt_device(bo.ptr<u64>(), bs.ptr<u64>(), bm.ptr<u64>(), batch, seed, (u32)kn, sigma, uniform_mask != 0, st)))
        return rc;
    return finish_all({&bs, &bm, &bo}, st);
}

// TLWE::decrypt for `batch` TLWEs under one secret key: phases, not yet decoded
int fhe_tlwe_decrypt(uint64_t kn, const uint64_t *sk, const uint64_t *ct, uint64_t *p, size_t batch) {
    if (batch == 0) return 0;
    FHE_REQUIRE(sk && ct && p, "fhe_tlwe_decrypt: null pointer");
    FHE_REQUIRE(kn >= 1 && kn < (1ull << 31), "fhe_tlwe_decrypt: kn out of range");
    cudaStream_t st = current_stream();
    IoBuf bs, bc, bo;
    int rc;
    if ((rc = bs.init(sk, kn * 8, true, false, st))) return rc;
    if ((rc = bc.init(ct, batch * (kn + 1) * 8, true, false, st))) return rc;
    if ((rc = bo.init(p, batch * 8, false, true, st))) return rc;
    if ((rc = tlwe_decrypt_device(bs.ptr<u64>(), bc.ptr<u64>(), bo.ptr<u64>(), batch, (u32)kn, st))) return rc;
    return finish_all({&bs, &bc, &bo}, st);
}

int fhe_tlwe_mod_switch(const uint64_t *ct, uint64_t q2, uint64_t *out, size_t len) {
    if (len == 0) return 0;
    FHE_REQUIRE(ct && out, "null pointer");
    FHE_REQUIRE(q2 >= 1 && (q2 & (q2 - 1)) == 0, "fhe_tlwe_mod_switch: q2 must be a power of two (torus.rs:58-66)");
    cudaStream_t st = current_stream();
    IoBuf bi, bo;
    int rc;
    if ((rc = bi.init(ct, len * 8, true, false, st))) return rc;
    if ((rc = bo.init(out, len * 8, false, true, st))) return rc;
    const u32 shift = 64 - (63 - __builtin_clzll(q2));
    if ((rc = shift_right_device(bi.ptr<u64>(), bo.ptr<u64>(), len, shift, st))) return rc;
    return finish_all({&bi, &bo}, st);
}

int fhe_sample_extract(uint64_t n, uint64_t k, const uint64_t *ct, uint64_t h, uint64_t *out, size_t batch) {
    if (batch == 0) return 0;
    FHE_REQUIRE(ct && out, "null pointer");
    FHE_REQUIRE(n >= 1 && k >= 1 && h < n, "fhe_sample_extract: need h < n");
    cudaStream_t st = current_stream();
    IoBuf bi, bo;
    int rc;
    if ((rc = bi.init(ct, batch * (k + 1) * n * 8, true, false, st))) return rc;
    if ((rc = bo.init(out, batch * (k * n + 1) * 8, false, true, st))) return rc;
    if ((rc = sample_extract_device(bi.ptr<u64>(), bo.ptr<u64>(), batch, (u32)n, (u32)k, (u32)h, st))) return rc;
    return finish_all({&bi, &bo}, st);
}

// compute_lookup_table (tlwe.rs:196-214): coefficients [0 x delta, 1 x delta, ..., (t-1) x delta] (delta = n/t) as an Rq
// mod t, encoded by TGLWE::encode (tglwe.rs:49-58: c * floor(u64::MAX / t), wrapping), mask = 0.
namespace {
__global__ void lookup_table_kernel(u64 *__restrict__ out, u64 n, u64 k, u64 t, u64 delta_n) {
    const u64 total = (k + 1) * n, delta = ~0ull / t;
    for (u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        u64 v = 0;
        if (i >= k * n) {
            const u64 c = i - k * n;
            if (c < t * delta_n) v = (c / delta_n) * delta;  // Zq::from_u64(t, v) = v for v < t
        }
        out[i] = v;
    }
}
}  // namespace
int fhe_compute_lookup_table(uint64_t n, uint64_t k, uint64_t t, uint64_t *table) {
    FHE_REQUIRE(table != nullptr, "fhe_compute_lookup_table: null pointer");
    FHE_REQUIRE(n >= 1 && k >= 1 && t >= 1 && t <= n, "fhe_compute_lookup_table: need 1 <= t <= n (delta = n/t >= 1)");
    cudaStream_t st = current_stream();
    IoBuf bo;
    int rc;
    if ((rc = bo.init(table, (k + 1) * n * 8, false, true, st))) return rc;
    lookup_table_kernel<<<(unsigned)(((k + 1) * n + 255) / 256), 256, 0, st>>>(bo.ptr<u64>(), n, k, t, n / t);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&bo}, st);
}

// blind_rotation (tlwe.rs:121-148).  bsk == NULL or as_written == 0: AS EXECUTED by the reference (the CMux
// closure is a lazy iterator that is dropped): acc = table.left_rotate(mod_switch(c).b).
// as_written != 0: additionally runs the loop the source spells out, for j in 1..k:
//   acc = cmux(bsk[j], acc, acc.left_rotate(mod_switch(c).a[j]))   (extension; no reference execution runs it).
int fhe_blind_rotate(uint64_t n, uint64_t k, const fhe_tggsw *const *bsk, int as_written, const uint64_t *table,
                     const uint64_t *ct, uint64_t c_kn, uint64_t *acc_out, size_t batch) {
    if (batch == 0) return 0;
    FHE_REQUIRE(table && ct && acc_out, "null pointer");
    FHE_REQUIRE(n >= 1 && k >= 1 && (n & (n - 1)) == 0, "fhe_blind_rotate: n must be a power of two");
    // T64::mod_switch asserts q2.is_power_of_two() (torus.rs:58-66); blind_rotation switches to q2 = k*n (tlwe.rs:125)
    FHE_REQUIRE(((k * n) & (k * n - 1)) == 0, "fhe_blind_rotate: k*n must be a power of two (T64::mod_switch, torus.rs:58-66)");
    FHE_REQUIRE(!as_written || k == 1 || bsk != nullptr, "fhe_blind_rotate: as_written needs the k TGGSW handles");
    cudaStream_t st = current_stream();
    const size_t glwe = (k + 1) * n;
    IoBuf bt, bc, bo;
    int rc;
    if ((rc = bt.init(table, glwe * 8, true, false, st))) return rc;
    if ((rc = bc.init(ct, batch * (c_kn + 1) * 8, true, false, st))) return rc;
    if ((rc = bo.init(acc_out, batch * glwe * 8, false, true, st))) return rc;
    {
        Scratch ext;
        if ((rc = ext.alloc(batch * (k * n + 1) * 8, st))) return rc;
        if ((rc = rotate_extract_device(bt.ptr<u64>(), bc.ptr<u64>(), ext.ptr<u64>(), bo.ptr<u64>(), batch, (u32)n, (u32)k,
                                        (u32)c_kn, st)))
            return rc;
    }
    if (as_written && k > 1) {
        FHE_REQUIRE(c_kn >= k, "fhe_blind_rotate: ciphertext has fewer than k mask elements");
        Scratch s_rot, s_nxt, s_hs;
        if ((rc = s_rot.alloc(batch * glwe * 8, st))) return rc;
        if ((rc = s_nxt.alloc(batch * glwe * 8, st))) return rc;
        if ((rc = s_hs.alloc(batch * 8, st))) return rc;
        u64 *rot = s_rot.ptr<u64>(), *nxt = s_nxt.ptr<u64>(), *hs = s_hs.ptr<u64>();
        const u32 shift = 64 - (63 - __builtin_clzll(k * n));
        for (u64 j = 1; j < k && !rc; j++) {
            FHE_REQUIRE(bsk[j] != nullptr && bsk[j]->n == n && bsk[j]->g.k == k, "fhe_blind_rotate: bad TGGSW handle");
            // hs[b] = mod_switch(c_b.a[j]) : strided gather + shift
            FHE_CUDA_OK(cudaMemcpy2DAsync(hs, 8, bc.ptr<u64>() + j, (c_kn + 1) * 8, 8, batch, cudaMemcpyDeviceToDevice, st));
            if ((rc = shift_right_device(hs, hs, batch, shift, st))) break;
            if ((rc = tn_left_rotate_device(bo.ptr<u64>(), rot, batch * (k + 1), (u32)n, hs, 0, (u32)(k + 1), st))) break;
            if ((rc = extprod_device(bsk[j]->g, bo.ptr<u64>(), rot, nxt, batch, st))) break;  // cmux(bsk[j], acc, rot)
            FHE_CUDA_OK(cudaMemcpyAsync(bo.ptr<u64>(), nxt, batch * glwe * 8, cudaMemcpyDeviceToDevice, st));
        }
        if (rc) return rc;
    }
    return finish_all({&bt, &bc, &bo}, st);
}

// bootstrapping (tlwe.rs:150-161) as executed: blind_rotation -> sample_extraction(0) -> key_switch(2, l, ksk).
int fhe_bootstrap(uint64_t n, uint64_t k, const fhe_ksk *ksk, const uint64_t *table, const uint64_t *ct, uint64_t c_kn,
                  uint64_t *out, size_t batch) {
    FHE_REQUIRE(ksk != nullptr, "null KSK handle");
    if (batch == 0) return 0;
    FHE_REQUIRE(table && ct && out, "null pointer");
    FHE_REQUIRE(n >= 1 && k >= 1 && (n & (n - 1)) == 0, "fhe_bootstrap: n must be a power of two");
    FHE_REQUIRE(((k * n) & (k * n - 1)) == 0, "fhe_bootstrap: k*n must be a power of two (T64::mod_switch, torus.rs:58-66)");
    FHE_REQUIRE(ksk->k.kn_in == k * n, "fhe_bootstrap: KSK input dimension must be k*n");
    cudaStream_t st = current_stream();
    IoBuf bt, bc, bo;
    int rc;
    if ((rc = bt.init(table, (k + 1) * n * 8, true, false, st))) return rc;
    const size_t chunk = 1024;  // ~8 MB of ciphertexts per stage; consecutive chunks compute on two streams
    if (batch >= 4 * chunk && is_host_ptr(ct) && is_host_ptr(out)) {
        // all-host call: H2D, rotate+extract+key switch and D2H of consecutive chunks overlap
        Scratch ext;
        const size_t ext_words = chunk * (k * n + 1);
        if ((rc = ext.alloc(2 * ext_words * 8, st))) return rc;
        const u64 *tab = bt.ptr<u64>();
        rc = run_host_pipelined(ct, (c_kn + 1) * 8, out, (ksk->k.kn_out + 1) * 8, batch, chunk, st,
                                [&](const void *din, void *dout, size_t nb, cudaStream_t s, int par) {
                                    u64 *e = ext.ptr<u64>() + (size_t)par * ext_words;
                                    int r = rotate_extract_device(tab, (const u64 *)din, e, nullptr, nb, (u32)n, (u32)k, (u32)c_kn, s);
                                    return r ? r : key_switch_device(ksk->k, e, (u64 *)dout, nb, s);
                                });
        return rc;
    }
    if ((rc = bc.init(ct, batch * (c_kn + 1) * 8, true, false, st))) return rc;
    if ((rc = bo.init(out, batch * (ksk->k.kn_out + 1) * 8, false, true, st))) return rc;
    Scratch ext;
    if ((rc = ext.alloc(batch * (k * n + 1) * 8, st))) return rc;
    rc = rotate_extract_device(bt.ptr<u64>(), bc.ptr<u64>(), ext.ptr<u64>(), nullptr, batch, (u32)n, (u32)k, (u32)c_kn, st);
    if (!rc) rc = key_switch_device(ksk->k, ext.ptr<u64>(), bo.ptr<u64>(), batch, st);
    if (rc) return rc;
    return finish_all({&bt, &bc, &bo}, st);
}

// Extension: blind rotation with one TGGSW per mask element, as a CMux chain kept on chip.
int fhe_bootstrap_chain(uint64_t n, uint64_t k, const fhe_tggsw *const *bsk, uint64_t steps, int mode, const fhe_ksk *ksk,
                        const uint64_t *table, const uint64_t *ct, uint64_t c_kn, uint64_t *out, size_t batch) {
    if (batch == 0) return 0;
    FHE_REQUIRE(table && ct && out && (steps == 0 || bsk), "fhe_bootstrap_chain: null pointer");
    FHE_REQUIRE(n >= 2 && k >= 1 && (n & (n - 1)) == 0, "fhe_bootstrap_chain: n must be a power of two");
    FHE_REQUIRE(mode != 0 || ((k * n) & (k * n - 1)) == 0,
                "fhe_bootstrap_chain: k*n must be a power of two in the as-written mode (T64::mod_switch, torus.rs:58-66)");
    FHE_REQUIRE(steps <= c_kn, "fhe_bootstrap_chain: more TGGSWs than mask elements");
    FHE_REQUIRE(ksk == nullptr || ksk->k.kn_in == k * n, "fhe_bootstrap_chain: KSK input dimension must be k*n");
    std::vector<const Tggsw *> gs(steps);
    for (uint64_t j = 0; j < steps; j++) {
        FHE_REQUIRE(bsk[j] != nullptr && bsk[j]->n == n && bsk[j]->g.k == k, "fhe_bootstrap_chain: bad TGGSW handle");
        gs[j] = &bsk[j]->g;
    }
    cudaStream_t st = current_stream();
    const size_t glwe = (k + 1) * n, kn = k * n, out_w = ksk ? ksk->k.kn_out + 1 : kn + 1;
    IoBuf bt, bc, bo;
    int rc;
    if ((rc = bt.init(table, glwe * 8, true, false, st))) return rc;
    if ((rc = bc.init(ct, batch * (c_kn + 1) * 8, true, false, st))) return rc;
    if ((rc = bo.init(out, batch * out_w * 8, false, true, st))) return rc;
    Scratch s_acc, s_acc2, s_hs, s_ext;
    if ((rc = s_acc.alloc(batch * glwe * 8, st))) return rc;
    if ((rc = s_acc2.alloc(batch * glwe * 8, st))) return rc;
    if ((rc = s_hs.alloc(batch * (steps ? steps : 1) * 8, st))) return rc;
    if (ksk && (rc = s_ext.alloc(batch * (kn + 1) * 8, st))) return rc;
    if ((rc = chain_prepare_device(bt.ptr<u64>(), bc.ptr<u64>(), s_acc.ptr<u64>(), s_hs.ptr<u64>(), batch, (u32)n, (u32)k,
                                   (u32)c_kn, (u32)steps, mode != 0, st)))
        return rc;
    const u64 *acc = s_acc.ptr<u64>();
    if (steps) {
        if ((rc = cmux_chain_device(gs.data(), steps, s_acc.ptr<u64>(), s_hs.ptr<u64>(), mode != 0, s_acc2.ptr<u64>(), batch, st)))
            return rc;
        acc = s_acc2.ptr<u64>();
    }
    u64 *ext = ksk ? s_ext.ptr<u64>() : bo.ptr<u64>();
    if ((rc = sample_extract_device(acc, ext, batch, (u32)n, (u32)k, 0, st))) return rc;
    if (ksk && (rc = key_switch_device(ksk->k, ext, bo.ptr<u64>(), batch, st))) return rc;
    return finish_all({&bt, &bc, &bo}, st);
}

}  // extern "C"
