/*
 * fhe_b200.h -- C ABI of libfhe_b200.so: the B200 (sm_100a) implementation of the ring-arithmetic hot
 * path of arnaucube/fhe-study.  This is the drop-in boundary: the reference has NO FFI layer today (all
 * of it is Rust operator overloading), so every entry point below names the reference function whose
 * body it replaces (paths relative to the reference tree); INTEGRATION.md shows the Rust `extern "C"`
 * block + `gpu`-feature call sites a maintainer would add.
 *
 * Conventions
 *   - every polynomial / ciphertext is a flat row-major array of uint64_t (SURVEY 8b layouts below);
 *   - EVERY data pointer may be a host pointer (pageable or pinned) or a device pointer; host buffers
 *     are staged through the device and the call returns when the results are back on the host;
 *     with device pointers only, the call is asynchronous on the stream set by fhe_set_stream();
 *   - all functions return 0 on success; on failure a negative code, with fhe_last_error() giving the
 *     reason (the Rust shim turns this into the panic!/anyhow! the reference raises at that site);
 *   - thread-safe: plan cache behind a lock, error string and stream selection are per thread.
 *
 * Layouts (uint64_t words)
 *   Rq / Tn polynomial : n coefficients
 *   TGLWE              : (k+1)*n   (mask polynomials 0..k-1, then the body)      tfhe/src/tggsw.rs:52-54
 *   TGLev              : l TGLWEs  (level j=0 first = gadget 2^63-1 = MSB digit)  tfhe/src/tggsw.rs:100-122
 *   TGGSW              : (k+1) TGLevs (rows for the k mask polynomials, body row last)
 *   TLWE               : kn+1      (mask, then b)
 *   KSK                : kn_in * l * (kn_out+1)                                   tfhe/src/tlwe.rs:84-100
 *   BFV RLWE / RLK     : 2*n       (c0 then c1 / rlk0 then rlk1)                  bfv/src/lib.rs:35-47
 */
#ifndef FHE_B200_H
#define FHE_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define FHE_API __attribute__((visibility("default")))
#else
#define FHE_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- runtime ------------------------------------------------------------------------------------ */
FHE_API const char *fhe_last_error(void);            /* per-thread message of the last failing call */
FHE_API int fhe_device_count(int *count);
FHE_API int fhe_set_device(int device);              /* cudaSetDevice for the calling thread */
FHE_API int fhe_current_device(int *device);         /* the calling thread's current device; every handle (plan, key) must
                                                        be used on the device it was created on, else the call fails */
FHE_API int fhe_set_stream(void *cuda_stream);       /* cudaStream_t used by this thread's calls (NULL = default) */
FHE_API int fhe_synchronize(void);                   /* wait for this thread's stream */
/* number of kernels this library has launched from this process (bench.py's gpu_launches) */
FHE_API uint64_t fhe_launch_count(void);

/* Register-only microbenchmarks of the integer pipes (the denominators of the integer rooflines):
 * kind 0 = 32-bit IMAD lane-ops/s, 1 = 32-bit Shoup modmul/s, 2 = 64-bit Shoup modmul/s (the better of the library's
 * chained form and the compiler's), 6 = 32-bit product in the Fermat form of q = 2^16 + 1 (an experiment);
 * kind 3 = int8 tensor-core ops/s (tcgen05.mma.kind::i8, M=128 N=256 K=32 issued back to back from shared-memory
 * operands, no loads) as a burst, 4 = the same sustained over ~2 s: the measured peaks the key-switch GEMM is quoted
 * against. */
FHE_API int fhe_int_peak(int kind, double *ops_per_s);

/* ---- NTT plans: arith/src/ntt.rs:18-38 (the (q,n) -> (roots, roots_inv, n_inv) cache) ------------- */
typedef struct fhe_ntt_plan fhe_ntt_plan;
/* Builds (or fetches from the per-device cache) the plan for Z_q[X]/(X^n+1).  Fails where the reference
 * panics (arith/src/ntt.rs:116-130): n not a power of two, 2n not dividing q-1, no primitive root; also
 * rejects q >= 2^63 (Zq::add would overflow, arith/src/zq.rs:225) and composite q. */
FHE_API int fhe_ntt_plan_create(uint64_t q, uint64_t n, fhe_ntt_plan **plan);
FHE_API void fhe_ntt_plan_destroy(fhe_ntt_plan *plan);
/* psi = the primitive 2n-th root the reference's search returns (arith/src/ntt.rs:115-131), n_inv, and
 * (optionally, may be NULL) the two n-entry tables in the reference's order roots[i] = psi^bitrev(i). */
FHE_API int fhe_ntt_plan_info(const fhe_ntt_plan *plan, uint64_t *psi, uint64_t *n_inv, uint64_t *roots, uint64_t *roots_inv);
/* Diagnostics: which kernel shape the plan selected -- config[0] = modular policy (0 Lazy32, 1 Lazy64, 2 Strict64,
 * 3 Small32, 4 Fermat32 = Small32 with radix-4 butterflies, q = 65537 only), [1] = log2(coefficients per thread), [2] = dual-operand polymul, [3] = NTT(a) parked in the output
 * row, [4] = persistent staged polymul.  (No reference counterpart; used by the tests and the tuning tools.) */
FHE_API int fhe_ntt_plan_config(const fhe_ntt_plan *plan, int *config);

/* PRECONDITION of every Rq entry point below: coefficients are canonical, 0 <= v < q, as the reference's Zq type
 * guarantees (arith/src/zq.rs:21-31).  The kernels do not reduce on load (the 32-bit policies read the low word
 * only), so unreduced words give wrong results silently; fhe_rq_check_canonical is the validation pass for callers
 * that cannot vouch for their buffers: *first_bad = index of the first word >= q, or UINT64_MAX if there is none. */
FHE_API int fhe_rq_check_canonical(uint64_t q, const uint64_t *words, size_t len, uint64_t *first_bad);

/* NTT::ntt (arith/src/ntt.rs:44-73): natural order in, bit-reversed order out; `batch` polynomials. */
FHE_API int fhe_ntt_fwd(const fhe_ntt_plan *plan, const uint64_t *in, uint64_t *out, size_t batch);
/* NTT::intt (arith/src/ntt.rs:78-110), including the n^-1 scaling. */
FHE_API int fhe_ntt_inv(const fhe_ntt_plan *plan, const uint64_t *in, uint64_t *out, size_t batch);

/* ring_nq::mul / mul_mut (arith/src/ring_nq.rs:564-607): c = intt(A . B) with A = a if
 * (flags & FHE_A_IS_EVALS) else ntt(a), same for b.  c_evals (may be NULL) receives A . B, the `evals`
 * the reference caches on the product (ring_nq.rs:606). */
#define FHE_A_IS_EVALS 1
#define FHE_B_IS_EVALS 2
#define FHE_B_BROADCAST 4 /* b is ONE polynomial (n words) multiplied into every a_i: GLWE * R (gfhe/src/glwe.rs:263-280) */
FHE_API int fhe_rq_mul(const fhe_ntt_plan *plan, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t batch, int flags,
               uint64_t *c_evals);
/* Packed 32-bit wire format of the three calls above for q <= 2^32 (the reference's only NTT modulus is 65537):
 * every polynomial is n uint32_t words.  Same values, half the PCIe bytes on the host-buffer path; the Rust shim
 * gathers Vec<Zq>.v (16-byte AoS, zq.rs:7-10) into either layout at the same cost.  Host or device pointers;
 * the call returns with the result complete. */
FHE_API int fhe_ntt_fwd_u32(const fhe_ntt_plan *plan, const uint32_t *in, uint32_t *out, size_t batch);
FHE_API int fhe_ntt_inv_u32(const fhe_ntt_plan *plan, const uint32_t *in, uint32_t *out, size_t batch);
FHE_API int fhe_rq_mul_u32(const fhe_ntt_plan *plan, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t batch, int flags,
                           uint32_t *c_evals);

/* Bit-packed wire format of the same three calls (q < 2^30, q <= 2^bits, 16 <= bits <= 31, n >= 1024): a polynomial is
 * n*bits/32 uint32_t words, coefficient i in bits [i*bits, (i+1)*bits) little-endian -- 17 bits per coefficient for
 * the reference's q = 65537 instead of 128 (Vec<Zq>), 64 or 32.  The kernels read and write this format directly
 * (one launch per chunk); it exists because the host-buffer path is PCIe-bound.  fhe_pack_bits / fhe_unpack_bits are
 * the host-side (de)serialisers (len % 32 == 0); the format is also the on-disk layout of fhe_b200_file.h. */
FHE_API int fhe_ntt_fwd_packed(const fhe_ntt_plan *plan, int bits, const uint32_t *in, uint32_t *out, size_t batch);
FHE_API int fhe_ntt_inv_packed(const fhe_ntt_plan *plan, int bits, const uint32_t *in, uint32_t *out, size_t batch);
FHE_API int fhe_rq_mul_packed(const fhe_ntt_plan *plan, int bits, const uint32_t *a, const uint32_t *b, uint32_t *c, size_t batch,
                              int flags, uint32_t *c_evals);
FHE_API int fhe_pack_bits(int bits, const uint64_t *in, uint32_t *out, size_t len);
FHE_API int fhe_unpack_bits(int bits, const uint32_t *in, uint64_t *out, size_t len);

/* ---- Tn = T_q[X]/(X^n+1), q = 2^64 (arith/src/ring_torus.rs) ------------------------------------------- */
/* impl Mul<Tn> for Tn -> naive_poly_mul (ring_torus.rs:251-298): exact negacyclic product mod 2^64. */
FHE_API int fhe_tn_mul(uint64_t n, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t batch);
/* Add / Sub / Neg (ring_torus.rs:153-249; also T64 vectors, torus.rs:80-153): wrapping, `len` words. */
FHE_API int fhe_tn_add(const uint64_t *a, const uint64_t *b, uint64_t *c, size_t len);
FHE_API int fhe_tn_sub(const uint64_t *a, const uint64_t *b, uint64_t *c, size_t len);
FHE_API int fhe_tn_neg(const uint64_t *a, uint64_t *c, size_t len);
/* Tn::left_rotate / TGLWE::left_rotate (ring_torus.rs:118-132, tfhe/src/tglwe.rs:116-119): multiply by
 * X^-h.  `polys` polynomials in groups of `group` consecutive polynomials (group = k+1 rotates whole TGLWEs);
 * group g uses h[g] (reduced mod n).  out must not alias a. */
FHE_API int fhe_tn_left_rotate(uint64_t n, const uint64_t *a, const uint64_t *h, uint64_t group, uint64_t *out, size_t polys);

/* ---- TGGSW external product / CMux (tfhe/src/tggsw.rs) ---------------------------------------------------- */
typedef struct fhe_tggsw fhe_tggsw;
/* Uploads one TGGSW ((k+1)*64 TGLWE rows, beta=2 / l=64 as hard-coded at tggsw.rs:49-50) and transforms it
 * once; the handle keeps it resident in HBM.  rows: (k+1) * 64 * (k+1) * n words. */
FHE_API int fhe_tggsw_load(uint64_t n, uint64_t k, const uint64_t *rows, fhe_tggsw **handle);
/* TGGSW::encrypt_s(sk, m) (tggsw.rs:17-33 over tggsw.rs:100-122 and glwe.rs:140-156, beta = 2, l = 64) ON THE DEVICE:
 * sk = k polynomials, m = the message polynomial (a bootstrapping key is k such calls with m = s_i, tlwe.rs:176-179).
 * Sampler and flags as fhe_ksk_generate (counter-based SplitMix64; the CPU restatement orc_tggsw_encrypt_s_ctr gives the
 * same rows bit for bit).  rows_out (host or device, may be NULL) receives the (k+1)*64*(k+1)*n sampled words.
 * SURVEY 8f rank 3. */
FHE_API int fhe_tggsw_generate(uint64_t n, uint64_t k, const uint64_t *sk, const uint64_t *m, double sigma, uint64_t seed,
                               int uniform_mask, uint64_t *rows_out, fhe_tggsw **handle);
FHE_API void fhe_tggsw_destroy(fhe_tggsw *handle);
/* impl Mul<TGLWE> for TGGSW (tggsw.rs:45-62): out_b = tggsw (x) ct_b for `batch` TGLWEs of (k+1)*n words. */
FHE_API int fhe_extprod(const fhe_tggsw *handle, const uint64_t *ct, uint64_t *out, size_t batch);
/* TGGSW::cmux (tggsw.rs:39-41): out_b = ct1_b + tggsw (x) (ct2_b - ct1_b). */
FHE_API int fhe_cmux(const fhe_tggsw *handle, const uint64_t *ct1, const uint64_t *ct2, uint64_t *out, size_t batch);

/* CMux chain -- the loop blind_rotation spells out (tlwe.rs:138-147) for `steps` TGGSWs, composed from
 * TGGSW::cmux (tggsw.rs:39-41) and TGLWE::left_rotate (tglwe.rs:116-119):
 *   acc_b <- cmux(bsk[j], acc_b, X^{-h[b*steps+j]} * acc_b)   for j = 0 .. steps-1.
 * negacyclic == 0: the rotation is left_rotate (h reduced mod n, ring_torus.rs:118-132);
 * negacyclic != 0: h is taken mod 2n and h >= n also negates (the true product by X^{-h}; the reference has no
 * such rotation, a working blind rotation needs it).  EXTENSION: no reference execution runs a
 * chain; parity is against the composition of the reference's own primitives.  When (n, k) has a fused kernel
 * the accumulator stays on chip for the whole chain (one HBM round trip per chain). */
FHE_API int fhe_cmux_chain(uint64_t n, uint64_t k, const fhe_tggsw *const *bsk, uint64_t steps, int negacyclic,
                           const uint64_t *acc_in, const uint64_t *h, uint64_t *acc_out, size_t batch);

/* ---- TLWE key switch / bootstrapping (tfhe/src/tlwe.rs, tfhe/src/tglwe.rs) ------------------------------ */
typedef struct fhe_ksk fhe_ksk;
/* KSK(Vec<TLev>) (tlwe.rs:84-100, tlev.rs:53-77): kn_in * l TLWE rows of kn_out+1 words, resident in HBM. */
FHE_API int fhe_ksk_load(uint64_t kn_in, uint64_t kn_out, uint64_t l, const uint64_t *rows, fhe_ksk **handle);
/* TLWE::new_ksk (tlwe.rs:84-100 over tlev.rs:53-77 and glwe.rs:140-156) ON THE DEVICE: row i*l + lv-1 encrypts
 * sk[i] * (u64::MAX / 2^lv) under new_sk.  The reference samples from an unseeded thread_rng that no implementation
 * can reproduce; the sampler here is a counter-based SplitMix64 (specified in oracle/fhe_oracle.c:
 * orc_tlwe_new_ksk_ctr, which the device output equals bit for bit): uniform_mask != 0 draws the mask uniformly from
 * Z_2^64, 0 draws it from the key distribution as the reference does (glwe.rs:146-149); error = round(sigma * x),
 * x an Irwin-Hall(12) approximation of N(0,1), cast like T64::rand (torus.rs:32-35).  SURVEY 8f rank 3. */
FHE_API int fhe_ksk_generate(uint64_t kn_in, uint64_t kn_out, uint64_t l, const uint64_t *sk, const uint64_t *new_sk,
                             double sigma, uint64_t seed, int uniform_mask, fhe_ksk **handle);
/* Reads the key rows back (layout of fhe_ksk_load); rows may be a host or a device pointer. */
FHE_API int fhe_ksk_export(const fhe_ksk *handle, uint64_t *rows);
FHE_API void fhe_ksk_destroy(fhe_ksk *handle);
/* TLWE::key_switch(param, 2, l, ksk) (tlwe.rs:101-112): `batch` TLWEs of kn_in+1 words -> kn_out+1 words. */
FHE_API int fhe_key_switch(const fhe_ksk *handle, const uint64_t *ct, uint64_t *out, size_t batch);
/* TLWE::encrypt_s (tlwe.rs:71-74 over glwe.rs:140-156) for `batch` already-encoded messages (TLWE::encode, tlwe.rs:52-59:
 * m * (u64::MAX / t)), sampled on the device with the counter-based sampler of fhe_ksk_generate: ciphertext b is row b of
 * the stream (the CPU restatement orc_tlwe_encrypt_ctr gives the same words). */
FHE_API int fhe_tlwe_encrypt(uint64_t kn, const uint64_t *sk, const uint64_t *msgs, double sigma, uint64_t seed,
                             int uniform_mask, uint64_t *ct, size_t batch);
/* TLWE::decrypt (tlwe.rs:80-82 over glwe.rs:175-179): p_b = ct_b.b - <ct_b.a, sk> for `batch` TLWEs of kn+1 words; the
 * phases are decoded with fhe_tn_mul_div_round(p, t, u64::MAX) and a reduction mod t (TLWE::decode, tlwe.rs:60-63). */
FHE_API int fhe_tlwe_decrypt(uint64_t kn, const uint64_t *sk, const uint64_t *ct, uint64_t *p, size_t batch);
/* TGLWE::encrypt_s (tglwe.rs:76-79 over glwe.rs:140-156) for `batch` already-encoded message polynomials (TGLWE::encode,
 * tglwe.rs:49-58), sampled on the device: ciphertext b uses the draws of row b of the stream of fhe_tggsw_generate (the CPU
 * restatement orc_tglwe_encrypt_ctr gives the same words). */
FHE_API int fhe_tglwe_encrypt(uint64_t n, uint64_t k, const uint64_t *sk, const uint64_t *msgs, double sigma, uint64_t seed,
                              int uniform_mask, uint64_t *ct, size_t batch);
/* TGLWE::decrypt (tglwe.rs:86-88 over glwe.rs:175-179): p_b = ct_b.b - sum_i ct_b.a_i * sk_i for `batch` TGLWEs; sk = k
 * polynomials, p = batch * n words (decode as above, TGLWE::decode tglwe.rs:59-63). */
FHE_API int fhe_tglwe_decrypt(uint64_t n, uint64_t k, const uint64_t *sk, const uint64_t *ct, uint64_t *p, size_t batch);
/* TLWE::mod_switch(q2) (tlwe.rs:114-118, torus.rs:58-66): every word >> (64 - log2 q2); q2 a power of two. */
FHE_API int fhe_tlwe_mod_switch(const uint64_t *ct, uint64_t q2, uint64_t *out, size_t len);
/* compute_lookup_table (tfhe/src/tlwe.rs:196-214 with TGLWE::encode, tglwe.rs:49-58): the trivial TGLWE (k zero mask
 * polynomials, then v) with v_c = floor(c / (n/t)) * floor(u64::MAX / t); (k+1)*n words.  n/t >= 1 and t * (n/t) <= n
 * as in the reference's parameters (a longer coefficient list would be folded by Rq::from_vec: not supported here). */
FHE_API int fhe_compute_lookup_table(uint64_t n, uint64_t k, uint64_t t, uint64_t *table);
/* TGLWE::sample_extraction(h) (tglwe.rs:89-115): `batch` TGLWEs ((k+1)*n) -> TLWEs (k*n+1). */
FHE_API int fhe_sample_extract(uint64_t n, uint64_t k, const uint64_t *ct, uint64_t h, uint64_t *out, size_t batch);
/* blind_rotation (tlwe.rs:121-148).  as_written == 0: AS THE REFERENCE EXECUTES IT (its CMux loop is a lazy
 * iterator that is dropped): acc = table.left_rotate(mod_switch(c, k*n).b).  as_written != 0 additionally runs
 * the loop the source spells out, for j in 1..k: acc = cmux(bsk[j], acc, acc.left_rotate(mod_switch(c).a[j]))
 * (an extension: no reference execution runs it).  bsk: k TGGSW handles (may be NULL when as_written == 0). */
FHE_API int fhe_blind_rotate(uint64_t n, uint64_t k, const fhe_tggsw *const *bsk, int as_written, const uint64_t *table,
                             const uint64_t *ct, uint64_t c_kn, uint64_t *acc_out, size_t batch);
/* bootstrapping (tlwe.rs:150-161) as executed: blind_rotation -> sample_extraction(0) -> key_switch.
 * table: one TGLWE ((k+1)*n words, e.g. compute_lookup_table's); ct: `batch` TLWEs of c_kn+1 words. */
FHE_API int fhe_bootstrap(uint64_t n, uint64_t k, const fhe_ksk *ksk, const uint64_t *table, const uint64_t *ct, uint64_t c_kn,
                          uint64_t *out, size_t batch);

/* Bootstrapping with one TGGSW per mask element (bsk: `steps` handles, steps <= c_kn) -- the blind rotation the
 * reference's loop is evidently meant to be (EXTENSION, see fhe_cmux_chain):
 *   mode 0 "as written": c' = mod_switch(c, k*n); acc = table.left_rotate(c'.b);
 *                        acc = cmux(bsk[j], acc, acc.left_rotate(c'.a[j])) for j < steps
 *   mode 1 "working":    c' = mod_switch(c, 2n); acc = X^{-c'.b} table; acc = cmux(bsk[j], acc, X^{+c'.a[j]} acc),
 *                        true negacyclic rotations, so acc = X^{-(b - <a,s>)} table when bsk[j] encrypts bit s_j
 * then sample_extraction(0) and, when ksk != NULL, key_switch (kn_in = k*n).  out: k*n+1 words per ciphertext
 * (ksk == NULL) or kn_out+1. */
FHE_API int fhe_bootstrap_chain(uint64_t n, uint64_t k, const fhe_tggsw *const *bsk, uint64_t steps, int mode,
                                const fhe_ksk *ksk, const uint64_t *table, const uint64_t *ct, uint64_t c_kn, uint64_t *out,
                                size_t batch);

/* ---- gfhe over Rq: GLev product and GLWE key switch (gfhe/src/glev.rs, gfhe/src/glwe.rs) -- SURVEY 8f rank 2 ------- */
typedef struct fhe_rq_glev fhe_rq_glev;
/* Uploads `rows` GLWE<Rq> ciphertexts ((k+1) polynomials of n words each: mask, then body) and keeps their NTT images
 * resident: a GLev (rows = l, glev.rs:14) or a KSK (rows = k*l, KSK[i][j] at row i*l + j; glwe.rs:99-125). */
FHE_API int fhe_rq_glev_load(const fhe_ntt_plan *plan, uint64_t k, uint64_t rows, const uint64_t *glwes, fhe_rq_glev **handle);
FHE_API void fhe_rq_glev_destroy(fhe_rq_glev *handle);
/* impl Mul<Vec<R>> for GLev<R> (glev.rs:67-80) with GLWE * R (glwe.rs:263-280): out_b = sum_r row_r * v_{b,r};
 * v: batch x rows polynomials, out: batch GLWEs. */
FHE_API int fhe_rq_glev_mul(const fhe_rq_glev *handle, const uint64_t *v, uint64_t *out, size_t batch);
/* GLWE<Rq>::key_switch(beta, l, ksk) (glwe.rs:126-137): (0, b) - sum_i ksk_i * a_i.decompose(beta, l), digits as
 * Zq::decompose (zq.rs:140-186, saturation branch included).  The handle must hold k*l rows. */
FHE_API int fhe_glwe_rq_key_switch(const fhe_rq_glev *ksk, uint32_t beta, uint32_t l, const uint64_t *ct, uint64_t *out,
                                   size_t batch);

/* ---- BFV ciphertext multiplication (bfv/src/lib.rs) ----------------------------------------------------------- */
/* RLWE::tensor (lib.rs:59-85): a, b = `batch` RLWEs (2n words) -> c0|c1|c2 (3n words each). */
FHE_API int fhe_bfv_tensor(uint64_t q, uint64_t n, uint64_t t, const uint64_t *a, const uint64_t *b, uint64_t *c012, size_t batch);
/* BFV::relinearize_204 (lib.rs:251-271): c0|c1|c2 and the RLK (2n words, coefficients mod pq) -> RLWE. */
FHE_API int fhe_bfv_relinearize(uint64_t q, uint64_t n, uint64_t pq, const uint64_t *rlk, const uint64_t *c012, uint64_t *out,
                                size_t batch);
/* RLWE::mul (lib.rs:87-90) = tensor + relinearize_204, fused. */
FHE_API int fhe_bfv_mul_relin(uint64_t q, uint64_t n, uint64_t t, uint64_t pq, const uint64_t *rlk, const uint64_t *a,
                              const uint64_t *b, uint64_t *out, size_t batch);
/* BFV::encrypt (lib.rs:142-160) for `batch` messages (n words mod t each) under the public key pk = (pk0, pk1) (2n words),
 * sampled on the device: c0 = pk0*u + e1 + m*floor(q/t), c1 = pk1*u + e2 with u from Uniform(-1,1) and e from a
 * Normal(0, sigma) stand-in, counter-based sampler (the CPU restatement orc_bfv_encrypt_ctr gives the same words). */
FHE_API int fhe_bfv_encrypt(const fhe_ntt_plan *plan, uint64_t q, uint64_t n, uint64_t t, const uint64_t *pk, const uint64_t *m,
                            double sigma, uint64_t seed, uint64_t *ct, size_t batch);
/* BFV::decrypt (lib.rs:164-178): m_b = ((c0 + c1 * s).mul_div_round(t, q)).remodule(t) for `batch` RLWEs (2n words each);
 * sk = the secret polynomial (n words), plan = the (q, n) plan (q and n are checked against it by the caller's types in the
 * reference; here they are passed so that the map kernels need no plan internals).  m: batch * n words. */
FHE_API int fhe_bfv_decrypt(const fhe_ntt_plan *plan, uint64_t q, uint64_t n, uint64_t t, const uint64_t *sk, const uint64_t *ct,
                            uint64_t *m, size_t batch);

/* BFV::new_key (lib.rs:120-140): s <- Uniform(0,2), a <- Uniform(0,q), e <- Normal(0, sigma); pk = (-a*s + e, a) with the
 * product through the (q, n) plan.  The reference samples from an unseeded thread_rng; the device uses the counter-based
 * sampler specified in oracle/fhe_oracle.c (orc_bfv_keygen_ctr), bit-exact against it.  sk: n words, pk: 2n words. */
FHE_API int fhe_bfv_keygen(const fhe_ntt_plan *plan, uint64_t q, uint64_t n, double sigma, uint64_t seed, uint64_t *sk, uint64_t *pk);
/* BFV::rlk_key (lib.rs:202-225) in the ring mod p*q, products through tmp_naive_mul (lib.rs:93-98):
 * rlk = (-(a*s + e) + (s*s)*p, a), 2n words mod p*q (p*q < 2^63).  Sampler: orc_bfv_rlk_key_ctr.  n <= 1024. */
FHE_API int fhe_bfv_rlk_generate(uint64_t q, uint64_t n, uint64_t p, double sigma, uint64_t seed, const uint64_t *sk, uint64_t *rlk);
/* BFV::mul_const (lib.rs:189-200): out_b = RLWE::mul(t, rlk, c_b, (m_b.remodule(q) * floor(q/t), 0)); m: batch * n words mod t. */
FHE_API int fhe_bfv_mul_const(uint64_t q, uint64_t n, uint64_t t, uint64_t pq, const uint64_t *rlk, const uint64_t *c,
                              const uint64_t *m, uint64_t *out, size_t batch);

/* ---- CKKS over Rq (ckks/src/lib.rs:46-119; the encoder of ckks/src/encoder.rs is out of scope: plaintexts are
 * elements of R = Z[X]/(X^n+1), int64 coefficients, as CKKS::encrypt / decrypt take and return them) ------------------ */
/* CKKS::new_key (lib.rs:46-63): s, a <- Uniform(-1,1) through Zq::from_f64, e <- Normal; pk = (-a*s + e, a). */
FHE_API int fhe_ckks_keygen(const fhe_ntt_plan *plan, uint64_t q, uint64_t n, double sigma, uint64_t seed, uint64_t *sk, uint64_t *pk);
/* CKKS::encrypt (lib.rs:66-84): ct_b = (m_b.to_rq(q) + e_0 + v*pk.0, v*pk.1 + e_1); m: batch * n int64, ct: batch * 2n words. */
FHE_API int fhe_ckks_encrypt(const fhe_ntt_plan *plan, uint64_t q, uint64_t n, const uint64_t *pk, const int64_t *m, double sigma,
                             uint64_t seed, uint64_t *ct, size_t batch);
/* CKKS::decrypt (lib.rs:86-94): m_b = (c.0 + c.1*s).mod_centered_q() (ring_n.rs:113-127), batch * n int64. */
FHE_API int fhe_ckks_decrypt(const fhe_ntt_plan *plan, uint64_t q, uint64_t n, const uint64_t *sk, const uint64_t *ct, int64_t *m,
                             size_t batch);
/* CKKS::add (lib.rs:113-115) and CKKS::sub (lib.rs:116-118; as written it ADDS the second components). */
FHE_API int fhe_ckks_add(uint64_t q, uint64_t n, const uint64_t *c0, const uint64_t *c1, uint64_t *out, size_t batch);
FHE_API int fhe_ckks_sub(uint64_t q, uint64_t n, const uint64_t *c0, const uint64_t *c1, uint64_t *out, size_t batch);

/* ---- coefficient-wise Rq / Tn operations (arith/src/ring_nq.rs, ring_torus.rs, zq.rs, torus.rs) ------------- */
/* Add / Sub / Neg / mul_by_u64 (ring_nq.rs:267-281,406-561) on `len` coefficients mod q (any q < 2^63). */
FHE_API int fhe_rq_add(uint64_t q, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t len);
FHE_API int fhe_rq_sub(uint64_t q, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t len);
FHE_API int fhe_rq_neg(uint64_t q, const uint64_t *a, uint64_t *c, size_t len);
FHE_API int fhe_rq_mul_u64(uint64_t q, const uint64_t *a, uint64_t s, uint64_t *c, size_t len);
/* Rq::remodule / mod_switch / mul_div_round (ring_nq.rs:82-113; f64 semantics of zq.rs:32-40,133-138). */
FHE_API int fhe_rq_remodule(const uint64_t *a, uint64_t p, uint64_t *c, size_t len);
FHE_API int fhe_rq_mod_switch(uint64_t q, const uint64_t *a, uint64_t p, uint64_t *c, size_t len);
FHE_API int fhe_rq_mul_div_round(uint64_t q, const uint64_t *a, uint64_t num, uint64_t den, uint64_t *c, size_t len);
/* Rq::from_vec_u64 (ring_nq.rs:55-63,132-141,156-159): reduce mod q and fold X^n = -1; in: `batch` x in_len. */
FHE_API int fhe_rq_from_vec(uint64_t q, uint64_t n, const uint64_t *in, uint64_t in_len, uint64_t *out, size_t batch);
/* Rq::decompose(beta, l) (ring_nq.rs:67-77, zq.rs:140-186): out = `polys` x l x n digit polynomials. */
FHE_API int fhe_rq_decompose(uint64_t q, uint64_t n, const uint64_t *a, uint32_t beta, uint32_t l, uint64_t *out, size_t polys);
/* Tn::decompose(2, l) (ring_torus.rs:67-77, torus.rs:43-52): out = `polys` x l x n bit polynomials. */
FHE_API int fhe_tn_decompose(uint64_t n, const uint64_t *a, uint32_t l, uint64_t *out, size_t polys);
/* Tn::mod_switch(p) -> Rq_p (ring_torus.rs:85-101), Mul<u64>/Mul<T64> (ring_torus.rs:300-327), mul_div_round. */
FHE_API int fhe_tn_mod_switch(const uint64_t *a, uint64_t p, uint64_t *c, size_t len);
FHE_API int fhe_tn_mul_u64(const uint64_t *a, uint64_t s, uint64_t *c, size_t len);
FHE_API int fhe_tn_mul_div_round(const uint64_t *a, uint64_t num, uint64_t den, uint64_t *c, size_t len);

#ifdef __cplusplus
}
#endif
#endif /* FHE_B200_H */
