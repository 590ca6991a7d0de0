// lib_torus.cu -- C-ABI entry points of the torus (q = 2^64) path: Tn arithmetic, TGGSW external product, CMux.
#include <map>
#include <memory>
#include <vector>

#include "../../include/fhe_b200.h"
#include "runtime.cuh"
#include "torus.cuh"

using namespace fhe;

namespace fhe {
static std::mutex g_tc_mu;
static std::map<std::pair<int, u64>, std::unique_ptr<TorusCtx>> g_tcs;

// per-(device, n) torus context, created on first use and kept for the life of the process
int get_torus_ctx(u64 n, TorusCtx **out) {
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tc_mu);
    auto key = std::make_pair(dev, n);
    auto it = g_tcs.find(key);
    if (it == g_tcs.end()) {
        std::unique_ptr<TorusCtx> tc(new TorusCtx());
        int rc = tc->init(n);
        if (rc) return rc;
        it = g_tcs.emplace(key, std::move(tc)).first;
    }
    *out = it->second.get();
    return 0;
}
}  // namespace fhe

extern "C" {

int fhe_tn_mul(uint64_t n, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t batch) {
    TorusCtx *tc;
    int rc = get_torus_ctx(n, &tc);
    if (rc) return rc;
    if (batch == 0) return 0;
    FHE_REQUIRE(a && b && c, "null polynomial pointer");
    cudaStream_t st = current_stream();
    const size_t bytes = batch * n * sizeof(u64);
    IoBuf ba, bb, bc;
    if ((rc = ba.init(a, bytes, true, false, st))) return rc;
    if ((rc = bb.init(b, bytes, true, false, st))) return rc;
    if ((rc = bc.init(c, bytes, false, true, st))) return rc;
    if ((rc = tn_mul_device(*tc, ba.ptr<u64>(), bb.ptr<u64>(), bc.ptr<u64>(), batch, st))) return rc;
    return finish_all({&ba, &bb, &bc}, st);
}

static int tn_elementwise(const uint64_t *a, const uint64_t *b, uint64_t *c, size_t len, int op) {
    if (len == 0) return 0;
    FHE_REQUIRE(a && c && (op == 2 || b), "null pointer");
    cudaStream_t st = current_stream();
    IoBuf ba, bb, bc;
    int rc;
    if ((rc = ba.init(a, len * 8, true, false, st))) return rc;
    if ((rc = bb.init(op == 2 ? nullptr : b, len * 8, true, false, st))) return rc;
    if ((rc = bc.init(c, len * 8, false, true, st))) return rc;
    if ((rc = tn_addsub_device(ba.ptr<u64>(), bb.ptr<u64>(), bc.ptr<u64>(), len, op, st))) return rc;
    return finish_all({&ba, &bb, &bc}, st);
}
int fhe_tn_add(const uint64_t *a, const uint64_t *b, uint64_t *c, size_t len) { return tn_elementwise(a, b, c, len, 0); }
int fhe_tn_sub(const uint64_t *a, const uint64_t *b, uint64_t *c, size_t len) { return tn_elementwise(a, b, c, len, 1); }
int fhe_tn_neg(const uint64_t *a, uint64_t *c, size_t len) { return tn_elementwise(a, nullptr, c, len, 2); }

int fhe_tn_left_rotate(uint64_t n, const uint64_t *a, const uint64_t *h, uint64_t group, uint64_t *out, size_t polys) {
    if (polys == 0) return 0;
    FHE_REQUIRE(a && h && out && a != out, "fhe_tn_left_rotate: null pointer or in-place call");
    FHE_REQUIRE(n >= 1 && group >= 1 && polys % group == 0, "fhe_tn_left_rotate: polys must be a multiple of group");
    cudaStream_t st = current_stream();
    IoBuf ba, bh, bo;
    int rc;
    if ((rc = ba.init(a, polys * n * 8, true, false, st))) return rc;
    if ((rc = bh.init(h, (polys / group) * 8, true, false, st))) return rc;
    if ((rc = bo.init(out, polys * n * 8, false, true, st))) return rc;
    if ((rc = tn_left_rotate_device(ba.ptr<u64>(), bo.ptr<u64>(), polys, (u32)n, bh.ptr<u64>(), 0, (u32)group, st)))
        return rc;
    return finish_all({&ba, &bh, &bo}, st);
}

int fhe_tggsw_load(uint64_t n, uint64_t k, const uint64_t *rows, fhe_tggsw **out) {
    FHE_REQUIRE(out != nullptr && rows != nullptr, "null pointer");
    *out = nullptr;
    FHE_REQUIRE(k >= 1 && k <= 64, "fhe_tggsw_load: k must be in 1..64");
    TorusCtx *tc;
    int rc = get_torus_ctx(n, &tc);
    if (rc) return rc;
    // exactness bound of the CRT lift: (k+1)*64*n*2^32 < P/2
    FHE_REQUIRE((unsigned __int128)(k + 1) * 64 * n * ((u64)1 << 32) < tc->cp.halfP,
                "fhe_tggsw_load: (k+1)*64*n too large for the exact two-prime lift");
    std::unique_ptr<fhe_tggsw> h(new fhe_tggsw());
    h->g.tc = tc;
    h->g.k = k;
    h->n = n;
    cudaStream_t st = current_stream();
    const size_t bytes = (k + 1) * 64 * (k + 1) * n * sizeof(u64);
    IoBuf br;
    if ((rc = br.init(rows, bytes, true, false, st))) return rc;
    if ((rc = tggsw_precompute(h->g, br.ptr<u64>(), st))) return rc;
    *out = h.release();
    return 0;
}
// TGGSW::encrypt_s on the device: the rows are sampled in HBM, transformed like loaded rows, and
// optionally copied out (rows_out: host or device, may be NULL).
int fhe_tggsw_generate(uint64_t n, uint64_t k, const uint64_t *sk, const uint64_t *m, double sigma, uint64_t seed,
                       int uniform_mask, uint64_t *rows_out, fhe_tggsw **out) {
    FHE_REQUIRE(out != nullptr && sk != nullptr && m != nullptr, "null pointer");
    *out = nullptr;
    FHE_REQUIRE(k >= 1 && k <= 64, "fhe_tggsw_generate: k must be in 1..64");
    TorusCtx *tc;
    int rc = get_torus_ctx(n, &tc);
    if (rc) return rc;
    FHE_REQUIRE((unsigned __int128)(k + 1) * 64 * n * ((u64)1 << 32) < tc->cp.halfP,
                "fhe_tggsw_generate: (k+1)*64*n too large for the exact two-prime lift");
    std::unique_ptr<fhe_tggsw> h(new fhe_tggsw());
    h->g.tc = tc;
    h->g.k = k;
    h->n = n;
    cudaStream_t st = current_stream();
    const size_t bytes = (k + 1) * 64 * (k + 1) * n * sizeof(u64);
    IoBuf bs, bm, bo;
    Scratch rows;
    if ((rc = bs.init(sk, k * n * 8, true, false, st))) return rc;
    if ((rc = bm.init(m, n * 8, true, false, st))) return rc;
    if ((rc = bo.init(rows_out, rows_out ? bytes : 0, false, true, st))) return rc;
    if ((rc = rows.alloc(bytes, st))) return rc;
    if ((rc = tggsw_generate_device(*tc, k, bs.ptr<u64>(), bm.ptr<u64>(), sigma, seed, uniform_mask != 0, rows.ptr<u64>(), st)))
        return rc;
    if (rows_out) FHE_CUDA_OK(cudaMemcpyAsync(bo.ptr<u64>(), rows.ptr<u64>(), bytes, cudaMemcpyDeviceToDevice, st));
    if ((rc = tggsw_precompute(h->g, rows.ptr<u64>(), st))) return rc;
    if ((rc = finish_all({&bs, &bm, &bo}, st))) return rc;
    FHE_CUDA_OK(cudaStreamSynchronize(st));
    *out = h.release();
    return 0;
}
// TGLWE::encrypt_s for `batch` already-encoded message polynomials (counter-based sampler)
int fhe_tglwe_encrypt(uint64_t n, uint64_t k, const uint64_t *sk, const uint64_t *msgs, double sigma, uint64_t seed,
                      int uniform_mask, uint64_t *ct, size_t batch) {
    if (batch == 0) return 0;
    FHE_REQUIRE(sk && msgs && ct, "fhe_tglwe_encrypt: null pointer");
    FHE_REQUIRE(k >= 1, "fhe_tglwe_encrypt: k must be >= 1");
    TorusCtx *tc;
    int rc = get_torus_ctx(n, &tc);
    if (rc) return rc;
    cudaStream_t st = current_stream();
    IoBuf bs, bm, bo;
    if ((rc = bs.init(sk, k * n * 8, true, false, st))) return rc;
    if ((rc = bm.init(msgs, batch * n * 8, true, false, st))) return rc;
    if ((rc = bo.init(ct, batch * (k + 1) * n * 8, false, true, st))) return rc;
    if ((rc = tglwe_encrypt_device(*tc, k, bs.ptr<u64>(), bm.ptr<u64>(), batch, sigma, seed, uniform_mask != 0, bo.ptr<u64>(), st)))
        return rc;
    return finish_all({&bs, &bm, &bo}, st);
}
// TGLWE::decrypt for `batch` TGLWEs under one secret key (k polynomials): phases, not yet decoded
int fhe_tglwe_decrypt(uint64_t n, uint64_t k, const uint64_t *sk, const uint64_t *ct, uint64_t *p, size_t batch) {
    if (batch == 0) return 0;
    FHE_REQUIRE(sk && ct && p, "fhe_tglwe_decrypt: null pointer");
    FHE_REQUIRE(k >= 1, "fhe_tglwe_decrypt: k must be >= 1");
    TorusCtx *tc;
    int rc = get_torus_ctx(n, &tc);
    if (rc) return rc;
    cudaStream_t st = current_stream();
    IoBuf bs, bc, bo;
    if ((rc = bs.init(sk, k * n * 8, true, false, st))) return rc;
    if ((rc = bc.init(ct, batch * (k + 1) * n * 8, true, false, st))) return rc;
    if ((rc = bo.init(p, batch * n * 8, false, true, st))) return rc;
    if ((rc = tglwe_decrypt_device(*tc, k, bs.ptr<u64>(), bc.ptr<u64>(), bo.ptr<u64>(), batch, st))) return rc;
    return finish_all({&bs, &bc, &bo}, st);
}
void fhe_tggsw_destroy(fhe_tggsw *h) {
    if (!h) return;
    cudaFree(h->g.R1);
    cudaFree(h->g.R2);
    cudaFree(h->g.R1f);
    cudaFree(h->g.R2f);
    delete h;
}

static int extprod_entry(const fhe_tggsw *h, const uint64_t *ct1, const uint64_t *ct2, uint64_t *out, size_t batch,
                         bool cmux) {
    FHE_REQUIRE(h != nullptr, "null TGGSW handle");
    if (batch == 0) return 0;
    FHE_REQUIRE(ct1 && out && (!cmux || ct2), "null ciphertext pointer");
    cudaStream_t st = current_stream();
    const size_t words = batch * (h->g.k + 1) * h->n, bytes = words * sizeof(u64);
    IoBuf b1, b2, bo;
    int rc;
    if ((rc = b1.init(ct1, bytes, true, false, st))) return rc;
    if ((rc = b2.init(cmux ? ct2 : nullptr, bytes, true, false, st))) return rc;
    if ((rc = bo.init(out, bytes, false, true, st))) return rc;
    rc = extprod_device(h->g, b1.ptr<u64>(), cmux ? b2.ptr<u64>() : nullptr, bo.ptr<u64>(), batch, st);
    if (rc) return rc;
    return finish_all({&b1, &b2, &bo}, st);
}
int fhe_extprod(const fhe_tggsw *h, const uint64_t *ct, uint64_t *out, size_t batch) {
    return extprod_entry(h, ct, nullptr, out, batch, false);
}
int fhe_cmux(const fhe_tggsw *h, const uint64_t *ct1, const uint64_t *ct2, uint64_t *out, size_t batch) {
    return extprod_entry(h, ct1, ct2, out, batch, true);
}


int fhe_cmux_chain(uint64_t n, uint64_t k, const fhe_tggsw *const *bsk, uint64_t steps, int negacyclic,
                   const uint64_t *acc_in, const uint64_t *h, uint64_t *acc_out, size_t batch) {
    if (batch == 0) return 0;
    FHE_REQUIRE(acc_in && acc_out && (steps == 0 || (bsk && h)), "fhe_cmux_chain: null pointer");
    FHE_REQUIRE(n >= 2 && (n & (n - 1)) == 0 && k >= 1, "fhe_cmux_chain: n must be a power of two");
    std::vector<const Tggsw *> gs(steps);
    for (uint64_t j = 0; j < steps; j++) {
        FHE_REQUIRE(bsk[j] != nullptr && bsk[j]->n == n && bsk[j]->g.k == k, "fhe_cmux_chain: bad TGGSW handle");
        gs[j] = &bsk[j]->g;
    }
    cudaStream_t st = current_stream();
    const size_t bytes = batch * (k + 1) * n * sizeof(u64);
    IoBuf bi, bh, bo;
    int rc;
    if ((rc = bi.init(acc_in, bytes, true, false, st))) return rc;
    if ((rc = bh.init(steps ? h : nullptr, batch * steps * sizeof(u64), true, false, st))) return rc;
    if ((rc = bo.init(acc_out, bytes, false, true, st))) return rc;
    if (steps == 0) {
        if (bo.ptr<u64>() != bi.ptr<u64>())
            FHE_CUDA_OK(cudaMemcpyAsync(bo.ptr<u64>(), bi.ptr<u64>(), bytes, cudaMemcpyDeviceToDevice, st));
    } else if ((rc = cmux_chain_device(gs.data(), steps, bi.ptr<u64>(), bh.ptr<u64>(), negacyclic != 0, bo.ptr<u64>(), batch,
                                       st))) {
        return rc;
    }
    return finish_all({&bi, &bh, &bo}, st);
}

}  // extern "C"
