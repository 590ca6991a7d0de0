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
FHE_API int fhe_set_stream(void *cuda_stream);       /* cudaStream_t used by this thread's calls (NULL = default) */
FHE_API int fhe_synchronize(void);                   /* wait for this thread's stream */
/* number of kernels this library has launched from this process (bench.py's gpu_launches) */
FHE_API uint64_t fhe_launch_count(void);

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

/* NTT::ntt (arith/src/ntt.rs:44-73): natural order in, bit-reversed order out; `batch` polynomials. */
FHE_API int fhe_ntt_fwd(const fhe_ntt_plan *plan, const uint64_t *in, uint64_t *out, size_t batch);
/* NTT::intt (arith/src/ntt.rs:78-110), including the n^-1 scaling. */
FHE_API int fhe_ntt_inv(const fhe_ntt_plan *plan, const uint64_t *in, uint64_t *out, size_t batch);

/* ring_nq::mul / mul_mut (arith/src/ring_nq.rs:564-607): c = intt(A . B) with A = a if
 * (flags & FHE_A_IS_EVALS) else ntt(a), same for b.  c_evals (may be NULL) receives A . B, the `evals`
 * the reference caches on the product (ring_nq.rs:606). */
#define FHE_A_IS_EVALS 1
#define FHE_B_IS_EVALS 2
FHE_API int fhe_rq_mul(const fhe_ntt_plan *plan, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t batch, int flags,
               uint64_t *c_evals);

#ifdef __cplusplus
}
#endif
#endif /* FHE_B200_H */
