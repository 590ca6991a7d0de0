/* fhe_b200_file.h -- flat, versioned on-disk / wire container for keys and ciphertexts of the hot path.
 *
 * The reference has no serialisation at all (no serde, SURVEY 5: "no on-disk format for keys or ciphertexts"); its
 * objects are nested Vec<Vec<..>> of 16-byte Zq{q,v} or 8-byte T64.  This container stores exactly the flat layouts the
 * C ABI takes (fhe_b200.h), so a file's payload can be handed to fhe_tggsw_load / fhe_ksk_load /
 * fhe_rq_mul* / fhe_bfv_* without any conversion -- and, for Rq data, in the same three word formats as the wire of
 * the host-buffer path (u64 words, u32 words, bit-packed).
 *
 * File = 96-byte little-endian header + payload.
 *   off  0  char[8]  magic "FHEB200\0"
 *   off  8  u32      version (1)
 *   off 12  u32      kind      (FHE_FILE_*)
 *   off 16  u32      encoding  (FHE_ENC_*)
 *   off 20  u32      bits      (coefficient width of FHE_ENC_PACKED, else 0)
 *   off 24  u64      q         (ring modulus; 0 = the torus 2^64)
 *   off 32  u64      n         (ring degree; 1 for scalar TLWE data)
 *   off 40  u64      k         (GLWE dimension, 0 if not applicable)
 *   off 48  u64      l         (gadget levels, 0 if not applicable)
 *   off 56  u64      count     (number of objects: polynomials, ciphertexts, ...)
 *   off 64  u64      words_per_object   (coefficients per object: n, (k+1)*n, kn+1, ...)
 *   off 72  u64      payload_bytes
 *   off 80  u64      checksum  (FNV-1a 64 over the payload bytes)
 *   off 88  u64      reserved (0)
 */
#ifndef FHE_B200_FILE_H
#define FHE_B200_FILE_H
#include <stddef.h>
#include <stdint.h>

#include "fhe_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

enum {
    FHE_FILE_RQ = 1,        /* Rq polynomials (arith/src/ring_nq.rs:19-27), n words each */
    FHE_FILE_TN = 2,        /* Tn polynomials (arith/src/ring_torus.rs:24-27) */
    FHE_FILE_TLWE = 3,      /* TLWE ciphertexts, kn+1 words (tfhe/src/tlwe.rs:37-40) */
    FHE_FILE_TGLWE = 4,     /* TGLWE ciphertexts, (k+1)*n words (tfhe/src/tglwe.rs:30) */
    FHE_FILE_TGGSW = 5,     /* TGGSW, (k+1)*l*(k+1)*n words (tfhe/src/tggsw.rs:14): input of fhe_tggsw_load */
    FHE_FILE_KSK = 6,       /* key-switching key, kn_in*l*(kn_out+1) words (tfhe/src/tlwe.rs:84-100): input of fhe_ksk_load;
                               n = kn_out, k = kn_in in the header */
    FHE_FILE_RLWE = 7,      /* BFV / CKKS ciphertexts, 2n words (bfv/src/lib.rs:35-36) */
    FHE_FILE_RLK = 8,       /* BFV relinearisation key, 2n words mod p*q (bfv/src/lib.rs:38) */
    FHE_FILE_SECRET = 9,    /* secret-key polynomial(s) / bit vectors */
    FHE_FILE_GLEV_RQ = 10   /* GLev<Rq> / GLWE<Rq> key-switching rows (gfhe/src/glev.rs): input of fhe_rq_glev_load */
};
enum { FHE_ENC_U64 = 0, FHE_ENC_U32 = 1, FHE_ENC_PACKED = 2 };

typedef struct fhe_file_info {
    uint32_t version, kind, encoding, bits;
    uint64_t q, n, k, l, count, words_per_object, payload_bytes, checksum;
} fhe_file_info;

/* payload size the header fields imply (0 on inconsistent fields, e.g. packed with words_per_object % 32 != 0) */
FHE_API uint64_t fhe_file_payload_bytes(const fhe_file_info *info);
/* Writes header + payload (host memory).  info->version / payload_bytes / checksum are filled in by the call. */
FHE_API int fhe_file_write(const char *path, fhe_file_info *info, const void *payload);
/* Reads and validates the header (magic, version, field consistency, file length). */
FHE_API int fhe_file_read_info(const char *path, fhe_file_info *info);
/* Reads the payload into `payload` (capacity bytes, host memory) and verifies the checksum. */
FHE_API int fhe_file_read_payload(const char *path, void *payload, size_t capacity);

#ifdef __cplusplus
}
#endif
#endif
