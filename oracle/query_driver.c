/*
 * query_driver.c — in-memory driver for the reference's own query executor.
 *
 * TEST / BASELINE INFRASTRUCTURE (see oracle/oracle_join.c for the rules).
 * This file is ours; it is compiled against the reference's headers where
 * they lie (/root/reference, never copied) and linked twice by
 * oracle/Makefile:
 *
 *   _ref/ref_driver   = this file + the UNMODIFIED reference objects
 *                       (query.o best_tree.o stats.o scheduler.o rhjoin.o
 *                       preprocess.o results.o filter.o inter_res.o) — the
 *                       reference's CPU path, used as parity oracle and as
 *                       the "reference" CPU baseline of bench.py;
 *   _ref/b200_driver  = this file + the reference's query.o best_tree.o
 *                       stats.o scheduler.o + libb200join.so in place of the
 *                       five operator objects — the link-time drop-in of
 *                       the reference's own ExecuteQuery
 *                       (query.c:325-467) calling the CUDA operators.
 *
 * It replaces handler.c:17-105 (stdin protocol) and relation_map.c:13-88
 * (mmap loader, limited to 2 GiB files by `int length`, relation_map.c:28)
 * by synthetic in-memory relations from include/b200_synth.h, or by relation
 * files read with fread.
 *
 * usage: driver [-t threads] [-r reps] [-w warm] [-T seconds] REL... -- 'query' ['query' ...]
 *   -T: time budget for the repetitions — stop once it is spent, but never before `warm` + 1 repetitions
 *       (bench.py's reference arm: the whole run must end within minutes whatever --steps asks for)
 *   REL := file:<path>
 *        | synth:<rows>:<col>,<col>,...   col := perm<k>[@seed] | pay[@seed]
 *                                              | zipf<k>[@seed] | uni<mod>[@seed] | iota
 * stdout: one result line per query and repetition (CalculateQueryResults)
 * stderr: one JSON line with wall-clock seconds per repetition
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "structs.h"
#include "query.h"
#include "scheduler.h"

#include "../include/b200_synth.h"

#ifdef B200_DROPIN
int  b200_init(int device);
int  b200_register_relations(const relation_map *map, int count);
int  b200_synchronize(void);
void b200_shutdown(void);
#endif

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* the statistics relation_map.c:53-83 computes: min, max, count, distinct
 * (boolean array over the value range, folded when the range is huge) */
static void column_stats_of(const uint64_t *col, uint64_t n, column_stats *st) {
    st->f = (double)n;
    st->l = st->u = n ? col[0] : 0;
    for (uint64_t k = 1; k < n; ++k) {
        if (col[k] > st->u) st->u = col[k];
        if (col[k] < st->l) st->l = col[k];
    }
    uint64_t size = st->u - st->l + 1;
    if (size > 50000000ull || size == 0) size = 50000000ull;
    unsigned char *seen = calloc(size, 1);
    uint64_t       d    = 0;
    for (uint64_t k = 0; k < n; ++k) {
        unsigned char *p = &seen[(col[k] - st->l) % size];
        d += !*p;
        *p = 1;
    }
    free(seen);
    st->d = (double)d;
}

static int parse_col(const char *spec, int *kind, uint64_t *k, uint64_t *seed, int col_index, int rel_index) {
    const char *at = strchr(spec, '@');
    *seed          = at ? strtoull(at + 1, NULL, 0) : (B200_SEED_R + 0x1000ull * (uint64_t)rel_index + (uint64_t)col_index);
    if (!strncmp(spec, "perm", 4)) { *kind = B200_SYNTH_PERM; *k = strtoull(spec + 4, NULL, 10); return 0; }
    if (!strncmp(spec, "pay", 3))  { *kind = B200_SYNTH_PAYLOAD; *k = 0; return 0; }
    if (!strncmp(spec, "zipf", 4)) { *kind = B200_SYNTH_ZIPF; *k = strtoull(spec + 4, NULL, 10); return 0; }
    if (!strncmp(spec, "uni", 3))  { *kind = B200_SYNTH_UNIFORM; *k = strtoull(spec + 3, NULL, 10); return 0; }
    if (!strncmp(spec, "iota", 4)) { *kind = B200_SYNTH_IOTA; *k = 0; return 0; }
    return 1;
}

static int load_relation(const char *spec, relation_map *rm, int rel_index) {
    if (!strncmp(spec, "file:", 5)) {
        FILE *f = fopen(spec + 5, "rb");
        if (!f) { perror(spec + 5); return 1; }
        uint64_t hdr[2];
        if (fread(hdr, 8, 2, f) != 2) return 1;
        rm->num_tuples  = hdr[0];
        rm->num_columns = hdr[1];
        rm->columns     = malloc(hdr[1] * sizeof(uint64_t *));
        rm->col_stats   = malloc(hdr[1] * sizeof(column_stats));
        for (uint64_t j = 0; j < hdr[1]; ++j) {
            rm->columns[j] = malloc((hdr[0] ? hdr[0] : 1) * 8);
            if (fread(rm->columns[j], 8, hdr[0], f) != hdr[0]) return 1;
            column_stats_of(rm->columns[j], hdr[0], &rm->col_stats[j]);
        }
        fclose(f);
        return 0;
    }
    if (!strncmp(spec, "synth:", 6)) {
        char *copy = strdup(spec + 6), *save = NULL;
        char *rows = strtok_r(copy, ":", &save);
        char *cols = strtok_r(NULL, ":", &save);
        if (!rows || !cols) return 1;
        rm->num_tuples = strtoull(rows, NULL, 0);
        int ncol = 1;
        for (char *p = cols; *p; ++p) ncol += *p == ',';
        rm->num_columns = (uint64_t)ncol;
        rm->columns     = malloc((size_t)ncol * sizeof(uint64_t *));
        rm->col_stats   = malloc((size_t)ncol * sizeof(column_stats));
        int   j     = 0;
        char *save2 = NULL;
        for (char *c = strtok_r(cols, ",", &save2); c; c = strtok_r(NULL, ",", &save2), ++j) {
            int      kind;
            uint64_t k, seed;
            if (parse_col(c, &kind, &k, &seed, j, rel_index)) { fprintf(stderr, "bad column spec %s\n", c); return 1; }
            uint64_t *col = malloc((rm->num_tuples ? rm->num_tuples : 1) * 8);
            for (uint64_t i = 0; i < rm->num_tuples; ++i) col[i] = b200_synth_value(kind, i, k, seed);
            rm->columns[j] = col;
            column_stats_of(col, rm->num_tuples, &rm->col_stats[j]);
        }
        free(copy);
        return 0;
    }
    fprintf(stderr, "bad relation spec %s\n", spec);
    return 1;
}

int main(int argc, char **argv) {
    int    threads = 4, reps = 1, warm = 0, a = 1;
    double budget  = 0.0;
    while (a < argc && argv[a][0] == '-' && strcmp(argv[a], "--")) {
        if (!strcmp(argv[a], "-t") && a + 1 < argc) threads = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "-r") && a + 1 < argc) reps = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "-w") && a + 1 < argc) warm = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "-T") && a + 1 < argc) budget = atof(argv[a + 1]);
        else { fprintf(stderr, "unknown option %s\n", argv[a]); return 2; }
        a += 2;
    }
    int first_rel = a, nrel = 0;
    while (a < argc && strcmp(argv[a], "--")) { ++a; ++nrel; }
    if (a >= argc - 1 || nrel == 0) {
        fprintf(stderr, "usage: %s [-t threads] [-r reps] REL... -- 'query' ...\n", argv[0]);
        return 2;
    }
    int first_query = a + 1, nquery = argc - first_query;

    double        t0      = now_s();
    relation_map *rel_map = calloc((size_t)nrel, sizeof(relation_map));
    for (int r = 0; r < nrel; ++r)
        if (load_relation(argv[first_rel + r], &rel_map[r], r)) return 2;
    double t_load = now_s() - t0;

    scheduler *sched = NULL;
    SchedulerInit(&sched, threads);   /* handler.c:61-63, thread count at run time */
#ifdef B200_DROPIN
    b200_init(-1);
    b200_register_relations(rel_map, nrel);   /* the hook after handler.c:52 */
    b200_synchronize();
#endif

    fprintf(stderr, "{\"threads\": %d, \"load_s\": %.3f, \"seconds\": [", threads, t_load);
    const double t_first = now_s();
    for (int rep = 0; rep < reps; ++rep) {
        if (budget > 0.0 && rep > warm && now_s() - t_first > budget) break;
        batch_listnode *batch = NULL;
        for (int q = 0; q < nquery; ++q) {
            char buff[250];
            snprintf(buff, sizeof buff, "%s\n", argv[first_query + q]);
            InsertToQueryBatch(&batch, buff);   /* handler.c:92 */
        }
        double t1 = now_s();
        for (batch_listnode *b = batch; b; b = b->next) ExecuteQuery(b, rel_map, sched);   /* handler.c:82-86 */
        fflush(stdout);
        double t2 = now_s();
        fprintf(stderr, "%s%.6f", rep ? ", " : "", t2 - t1);
        FreeBatch(batch);
    }
    fprintf(stderr, "]}\n");
    SchedulerDestroy(sched);
#ifdef B200_DROPIN
    b200_shutdown();
#endif
    return 0;
}
