/*
 * oracle_join.c — CPU restatement of the join hot path of VagelisN/Sigmod-2018.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sigmod-2018_b200/ may link, import
 * or execute this file; it is used by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py, and only as the checker.
 *
 * Parity pin: tests/test_oracle_vs_reference.py runs every function below
 * against the reference's own compiled objects (oracle/_ref/libref_ops.so,
 * built from /root/reference by oracle/Makefile) on seeded inputs, including
 * exact output ORDER, and tests/golden/ holds vectors generated from those
 * objects (tests/golden/make_golden.py) so the pin travels to boxes that have
 * no /root/reference.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference).  This is a restatement in flat arrays: the reference's
 * linked result lists, AoS tuples and pthread jobs are replaced by plain
 * buffers and loops that produce the same values in the same order.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b200_synth.h"

#define ORC_EXPORT __attribute__((visibility("default")))

/* rhjoin.c:311-325 HashFunction1: the n low bits of num */
ORC_EXPORT uint64_t orc_hash1(uint64_t num, uint64_t n) {
    uint64_t mask = ~0ull;
    if (n == 0) return 0; /* the reference is only ever called with N_LSB >= 1 */
    mask = mask << (64 - n);
    mask = mask >> (64 - n);
    return num & mask;
}

/* rhjoin.c:327-346 FindNextPrime, including its `i*i < num` bound (squares of
 * primes pass as "prime") and the `int i` loop counter */
ORC_EXPORT uint64_t orc_next_prime(uint64_t num) {
    if (num % 2 == 0) num++;
    for (;;) {
        int found = 1;
        for (int i = 3; (uint64_t)((int64_t)i * i) < num; ++i) {
            if (num % (uint64_t)i == 0) {
                found = 0;
                num += 2;
                break;
            }
        }
        if (found) break;
    }
    return num;
}

/* rhjoin.c:348-351 HashFunction2 */
ORC_EXPORT uint64_t orc_hash2(uint64_t num, uint64_t prime) { return num % prime; }

/*
 * preprocess.c:13-178 ReorderArray (parallel) == preprocess.c:302-362
 * SerialReorderArray in its result: histogram of key & (2^n_lsb - 1)
 * (HistJob 181-195), psum = running start with -1 for an empty bucket
 * (83-102), and a STABLE partition-contiguous copy (PartitionJob 222-299 fills
 * output ranges in input order; the serial variant is a counting sort).
 * rids == NULL means row_id = position (inter_res.c:202, 225).
 * Returns 0, or 1 when the input is empty (every bucket empty =>
 * preprocess.c:103-108 hands back NULL).
 */
ORC_EXPORT int orc_reorder(const uint64_t *keys, const uint64_t *rids, uint64_t n, int n_lsb,
                           uint64_t *out_keys, uint64_t *out_rids, uint64_t *hist, int64_t *psum) {
    const uint64_t nb = 1ull << n_lsb;
    for (uint64_t b = 0; b < nb; ++b) {
        hist[b] = 0;
        psum[b] = -1;
    }
    for (uint64_t i = 0; i < n; ++i) hist[orc_hash1(keys[i], (uint64_t)n_lsb)]++;
    int64_t  start = 0;
    int64_t *cur   = (int64_t *)malloc(nb * sizeof(int64_t));
    for (uint64_t b = 0; b < nb; ++b) {
        cur[b] = -1;
        if (hist[b] > 0) {
            psum[b] = start;
            cur[b]  = start;
            start += (int64_t)hist[b];
        }
    }
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t b = orc_hash1(keys[i], (uint64_t)n_lsb);
        const int64_t  p = cur[b]++;
        out_keys[p]      = keys[i];
        out_rids[p]      = rids ? rids[i] : i;
    }
    free(cur);
    return n == 0;
}

/*
 * rhjoin.c:253-273 InitIndex + rhjoin.c:219-250 CreateIndex on one bucket of
 * `cnt` keys: bucket[] has next_prime(cnt) slots of -1, chain[] cnt slots;
 * tuples are inserted from the last to the first, the first hit of a slot
 * becomes the head (1-based position), later hits are appended to the chain
 * TAIL, so a chain lists positions in descending order; 0 terminates.
 */
static void orc_build_index(const uint64_t *keys, uint64_t cnt, uint64_t prime, int64_t *bucket,
                            int64_t *chain) {
    for (uint64_t i = 0; i < prime; ++i) bucket[i] = -1;
    for (uint64_t i = 0; i < cnt; ++i) chain[i] = -1;
    for (int64_t i = (int64_t)cnt - 1; i >= 0; --i) {
        const uint64_t h = orc_hash2(keys[i], prime);
        if (bucket[h] == -1) {
            bucket[h] = i + 1;
            chain[i]  = 0;
        } else {
            int64_t sp = bucket[h] - 1;
            while (chain[sp] != 0) sp = chain[sp] - 1;
            chain[sp] = i + 1;
            chain[i]  = 0;
        }
    }
}

/* exported for the index-level parity test (bucket/chain contents) */
ORC_EXPORT uint64_t orc_create_index(const uint64_t *keys, uint64_t cnt, int64_t *bucket, int64_t *chain) {
    const uint64_t prime = orc_next_prime(cnt);
    if (bucket && chain) orc_build_index(keys, cnt, prime, bucket, chain);
    return prime;
}

/*
 * rhjoin.c:13-111 RadixHashJoin with rhjoin.c:113-137 JoinJob, 141-217
 * GetResults and 354-392 MergeResults: for every bucket (ascending) where
 * both sides are non-empty, index the side with FEWER tuples (S when
 * histR >= histS) and probe with the other in its reordered order; a probe
 * tuple emits the head match first, then the chain.  Pairs are always
 * (row_idR, row_idS).  The concatenation over buckets is what MergeResults
 * produces.  Writes at most `cap` pairs, returns the total number of pairs;
 * returns UINT64_MAX for the reference's NULL result (an empty input,
 * rhjoin.c:15-16).
 */
ORC_EXPORT uint64_t orc_radix_hash_join(const uint64_t *keys_r, const uint64_t *rids_r, uint64_t n_r,
                                        const uint64_t *keys_s, const uint64_t *rids_s, uint64_t n_s,
                                        int n_lsb, uint64_t *out_r, uint64_t *out_s, uint64_t cap) {
    if (n_r == 0 || n_s == 0) return UINT64_MAX;
    const uint64_t nb = 1ull << n_lsb;
    uint64_t *kr = malloc(n_r * 8), *rr = malloc(n_r * 8), *ks = malloc(n_s * 8), *rs = malloc(n_s * 8);
    uint64_t *hr = malloc(nb * 8), *hs = malloc(nb * 8);
    int64_t  *pr = malloc(nb * 8), *ps = malloc(nb * 8);
    orc_reorder(keys_r, rids_r, n_r, n_lsb, kr, rr, hr, pr);
    orc_reorder(keys_s, rids_s, n_s, n_lsb, ks, rs, hs, ps);
    uint64_t m = 0;
    for (uint64_t b = 0; b < nb; ++b) {
        if (hr[b] == 0 || hs[b] == 0) continue;
        /* r_s == 0: S indexed, R probes (rhjoin.c:118-125); else R indexed */
        const int       index_s = hr[b] >= hs[b];
        const uint64_t *ik = index_s ? ks + ps[b] : kr + pr[b];
        const uint64_t *ir = index_s ? rs + ps[b] : rr + pr[b];
        const uint64_t  ic = index_s ? hs[b] : hr[b];
        const uint64_t *fk = index_s ? kr + pr[b] : ks + ps[b];
        const uint64_t *fr = index_s ? rr + pr[b] : rs + ps[b];
        const uint64_t  fc = index_s ? hr[b] : hs[b];
        const uint64_t  prime  = orc_next_prime(ic);
        int64_t        *bucket = malloc(prime * 8), *chain = malloc(ic * 8);
        orc_build_index(ik, ic, prime, bucket, chain);
        for (uint64_t i = 0; i < fc; ++i) {
            const uint64_t h = orc_hash2(fk[i], prime);
            if (bucket[h] == -1) continue;
            int64_t pos = bucket[h]; /* 1-based */
            while (pos != 0) {
                if (ik[pos - 1] == fk[i]) {
                    if (m < cap) {
                        out_r[m] = index_s ? fr[i] : ir[pos - 1];
                        out_s[m] = index_s ? ir[pos - 1] : fr[i];
                    }
                    ++m;
                }
                pos = chain[pos - 1];
            }
        }
        free(bucket);
        free(chain);
    }
    free(kr); free(rr); free(ks); free(rs); free(hr); free(hs); free(pr); free(ps);
    return m;
}

/*
 * filter.c:92-190 Filter: `col[i] cmp (uint64)(int)value` over the base column
 * (ids == NULL; emits base row ids, filter.c:113-122) or through the
 * intermediate's row ids (emits POSITIONS, filter.c:124-133).  The constant is
 * an `int` (structs.h:146) promoted by the usual arithmetic conversions.
 * Returns the number of hits; 0 is the reference's NULL result.
 */
ORC_EXPORT uint64_t orc_filter(const uint64_t *col, uint64_t n, const uint64_t *ids, uint64_t n_ids,
                               char cmp, int value, uint64_t *out) {
    const uint64_t c     = (uint64_t)(int64_t)value;
    const uint64_t count = ids ? n_ids : n;
    uint64_t       m     = 0;
    for (uint64_t i = 0; i < count; ++i) {
        const uint64_t v = ids ? col[ids[i]] : col[i];
        const int keep   = cmp == '>' ? v > c : cmp == '<' ? v < c : v == c;
        if (keep) out[m++] = i;
    }
    return m;
}

/* inter_res.c:79-98, 119-137; filter.c:60-76; inter_res.c:304-313 — gather of
 * one row-id column through a list of positions */
ORC_EXPORT void orc_gather(const uint64_t *in, const uint64_t *pos, uint64_t m, uint64_t *out) {
    for (uint64_t i = 0; i < m; ++i) out[i] = in[pos[i]];
}

/* inter_res.c:376-385 JoinInterNode predicate: positions p with
 * colA[idA[p]] == colB[idB[p]] */
ORC_EXPORT uint64_t orc_inter_equal(const uint64_t *col_a, const uint64_t *id_a, const uint64_t *col_b,
                                    const uint64_t *id_b, uint64_t n, uint64_t *out) {
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; ++i)
        if (col_a[id_a ? id_a[i] : i] == col_b[id_b ? id_b[i] : i]) out[m++] = i;
    return m;
}

/* inter_res.c:320-339 CalculateQueryResults, one projection: sum of
 * col[ids[j]] mod 2^64 (ids == NULL: identity) */
ORC_EXPORT uint64_t orc_checksum(const uint64_t *col, const uint64_t *ids, uint64_t m) {
    uint64_t s = 0;
    for (uint64_t j = 0; j < m; ++j) s += col[ids ? ids[j] : j];
    return s;
}

/*
 * The 2-relation query `0 1|0.a=1.b|...` end to end, as query.c:408-463 runs
 * it: GetRelation x2 (inter_res.c:208-231), RadixHashJoin,
 * InsertJoinToInterResults on an empty node (inter_res.c:39-62: split the
 * pairs into two row-id columns), CalculateQueryResults.  proj_side[k] = 0
 * sums proj[k] through the R row ids, 1 through the S row ids.
 * Returns the number of result rows.
 */
ORC_EXPORT uint64_t orc_join_sum(const uint64_t *keys_r, uint64_t n_r, const uint64_t *keys_s, uint64_t n_s,
                                 int n_lsb, int n_proj, const uint64_t *const *proj, const int *proj_side,
                                 uint64_t *out_sums) {
    for (int k = 0; k < n_proj; ++k) out_sums[k] = 0;
    if (n_r == 0 || n_s == 0) return 0;
    uint64_t m = orc_radix_hash_join(keys_r, NULL, n_r, keys_s, NULL, n_s, n_lsb, NULL, NULL, 0);
    uint64_t *idr = malloc((m ? m : 1) * 8), *ids = malloc((m ? m : 1) * 8);
    orc_radix_hash_join(keys_r, NULL, n_r, keys_s, NULL, n_s, n_lsb, idr, ids, m);
    for (int k = 0; k < n_proj; ++k) out_sums[k] = orc_checksum(proj[k], proj_side[k] == 0 ? idr : ids, m);
    free(idr);
    free(ids);
    return m;
}

/* synthetic columns (include/b200_synth.h) */
ORC_EXPORT void orc_synth_column(uint64_t *out, uint64_t first, uint64_t n, int kind, uint64_t k,
                                 uint64_t seed) {
    for (uint64_t i = 0; i < n; ++i) out[i] = b200_synth_value(kind, first + i, k, seed);
}

/* relation_map.c:53-83 — the column statistics the reference's loader keeps: l = smallest value, u = largest,
 * d = "distinct values" as the reference counts them: a marker array of min(u - l + 1, 50 000 000) entries, entry
 * v - l when that size is below 50 000 000, else entry (v - l) % 5 000 000 (sic, relation_map.c:71), then the
 * number of marked entries (relation_map.c:75-79).  n >= 1 (the reference reads row 0 unconditionally). */
ORC_EXPORT void orc_column_stats(const uint64_t *col, uint64_t n, uint64_t *out_l, uint64_t *out_u, uint64_t *out_d) {
    uint64_t l = col[0], u = col[0];
    for (uint64_t k = 1; k < n; ++k) {
        if (col[k] > u) u = col[k];
        if (col[k] < l) l = col[k];
    }
    uint64_t size = u - l + 1;
    if (size > 50000000ull) size = 50000000ull;
    unsigned short *seen = calloc(size, sizeof(unsigned short));
    for (uint64_t k = 0; k < n; ++k) {
        if (size < 50000000ull) seen[col[k] - l] = 1;
        else seen[(col[k] - l) % 5000000ull] = 1;
    }
    uint64_t d = 0;
    for (uint64_t k = 0; k < size; ++k) d += seen[k] == 1;
    free(seen);
    *out_l = l;
    *out_u = u;
    *out_d = d;
}
