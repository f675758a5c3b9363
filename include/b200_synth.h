/*
 * b200_synth.h — the synthetic relation generator of BASELINE.json's configs
 * (SURVEY §8d).  One definition, integer-only, so that the CUDA generator in
 * libb200join.so, the CPU baseline driver (oracle/ref_driver.c) and the tests
 * produce bit-identical columns from (row index, seed) alone.
 *
 *   b200_perm(x, k, seed)   seed-keyed bijection on k-bit integers: unique,
 *                           pseudo-random in every radix bit (key columns).
 *   b200_payload(i, seed)   splitmix64(seed + i) & 0xFFFFFF (payload columns).
 *   b200_zipf_rank(i, k, seed)  rank in [0, 2^k): the octave e is uniform in
 *                           [0,k] and the rank uniform inside [2^e-1, 2^(e+1)-1),
 *                           i.e. P(rank = z) ~ 1/(z+1) — Zipf with theta = 1.0
 *                           (every octave of ranks carries the same mass).
 *
 * Not part of the reference (it ships only the `small` workload); the shapes
 * are the ones BASELINE.json names.
 */
#ifndef B200_SYNTH_H
#define B200_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define B200_HD __host__ __device__ __forceinline__
#else
#define B200_HD static inline
#endif

#define B200_SEED_R 0x51670D180001ull
#define B200_SEED_S 0x51670D180002ull

B200_HD uint64_t b200_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* every step is invertible mod 2^k: add, multiply by an odd constant,
 * xor with a right shift of at least k/2 */
B200_HD uint64_t b200_perm(uint64_t x, int k, uint64_t seed) {
    const uint64_t mask = (k >= 64) ? ~0ull : ((1ull << k) - 1ull);
    const int      s    = (k + 1) / 2;
    x = (x + seed) & mask;
    x = (x * 0x9E3779B97F4A7C15ull) & mask;
    x ^= x >> s;
    x = (x * 0xBF58476D1CE4E5B9ull) & mask;
    x ^= x >> s;
    x = (x * 0x94D049BB133111EBull) & mask;
    x ^= x >> s;
    return x;
}

B200_HD uint64_t b200_payload(uint64_t i, uint64_t seed) {
    return b200_splitmix64(seed + i) & 0xFFFFFFull;
}

B200_HD uint64_t b200_zipf_rank(uint64_t i, int k, uint64_t seed) {
    const uint64_t u    = b200_splitmix64(seed ^ (i * 0xD6E8FEB86659FD93ull));
    const uint64_t mask = (1ull << k) - 1ull;
    /* octave e in [0, k): ranks [2^e - 1, 2^(e+1) - 1) */
    const int      e    = (int)((u >> 40) % (uint64_t)k);
    const uint64_t r    = (u & 0xFFFFFFFFFFull) & ((1ull << e) - 1ull);
    return (((1ull << e) - 1ull) + r) & mask;
}

/* column kinds understood by b200_synth_column / orc_synth_column */
enum {
    B200_SYNTH_PERM    = 0, /* col[i] = perm_k(i, seed)                      */
    B200_SYNTH_PAYLOAD = 1, /* col[i] = splitmix64(seed + i) & 0xFFFFFF      */
    B200_SYNTH_ZIPF    = 2, /* col[i] = perm_k(zipf_rank(i), seed2 = seed^1) */
    B200_SYNTH_UNIFORM = 3, /* col[i] = splitmix64(seed + i) % 2^k ... see   */
    B200_SYNTH_IOTA    = 4  /* col[i] = i                                    */
};

/* value of row i of a synthetic column; `k` is the bit width (PERM, ZIPF) or
 * the modulus (UNIFORM: values in [0, k)). */
B200_HD uint64_t b200_synth_value(int kind, uint64_t i, uint64_t k, uint64_t seed) {
    switch (kind) {
        case B200_SYNTH_PERM:    return b200_perm(i, (int)k, seed);
        case B200_SYNTH_PAYLOAD: return b200_payload(i, seed);
        case B200_SYNTH_ZIPF:    return b200_perm(b200_zipf_rank(i, (int)k, seed), (int)k, B200_SEED_R);
        case B200_SYNTH_UNIFORM: return b200_splitmix64(seed + i) % (k ? k : 1);
        default:                 return i;
    }
}

#endif /* B200_SYNTH_H */
