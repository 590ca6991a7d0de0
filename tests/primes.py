"""NTT-friendly test primes (2^16 | q-1), one per modular policy of the CUDA library."""
Q17 = 2**16 + 1                    # the reference's only NTT modulus (F5), Lazy32
Q30 = 0x3FFC0001                   # largest prime < 2^30 with 2^16 | q-1, Lazy32
Q62 = 0x3FFFFFFFFFFF0001           # largest prime < 2^62 with 2^16 | q-1, Lazy64
Q63 = 0x7FFFFFFFFFEF0001           # largest prime < 2^63 with 2^16 | q-1, Strict64 (reference limit q < 2^63)
Q22 = 0x390001                     # largest prime < 2^22 with 2^16 | q-1, Small32 (relaxed-lazy policy)
