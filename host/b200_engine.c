/*
 * b200_engine.c — a stand-alone C host for libb200join.so with the reference's process protocol.
 *
 * The reference's host side (handler.c, query.c, best_tree.c, stats.c, relation_map.c, scheduler.c) links
 * against the library unchanged.  This file is the same thing written from scratch, for
 * boxes that do not have the reference: it speaks the contest protocol of handler.c:17-105 — relation file
 * names until `Done`, then query lines, `F` ends a batch, one result line per query — loads relations the
 * way relation_map.c:13-88 does (mmap, header [rows][cols], column-major uint64; files larger than 2 GiB
 * are fine here), parses queries with the semantics of query.c:44-249, and runs them through the operator
 * API in ExecuteQuery's order (query.c:325-467).
 *
 * Two things differ from the reference on purpose:
 *   - scheduler (scheduler.c:9-132): the reference's jobs are slices of ONE join and a batch runs query by
 *     query (handler.c:78-89).  Here a job is a whole query: a batch is handed to a pool of worker threads,
 *     each of which owns a CUDA stream inside the library (thread-local context), and the result lines are
 *     printed in submission order when the batch ends — config 5's "concurrent queries on GPU streams".
 *   - join order: textual order with one rule, "start from a filtered binding if there is one"; the
 *     reference's JoinEnum DP (best_tree.c:105-223) only changes cost, never results.
 *
 * usage: b200_engine [-w workers] [-g gpus]      (defaults 4 and 1; B200_WORKERS / B200_GPUS override)
 *
 * -g N (N = 2..8): every relation is additionally position-sharded over N GPUs when it is loaded, and a query that is
 * one equi-join of two relations with at most one SUM per side and no filter (BASELINE config 2's shape) runs through
 * the library's multi-GPU plan (b200_join_sum_multi: one host thread per GPU, no NCCL); every other query runs on
 * GPU 0 as before.
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "../include/b200_join.h"

#define MAX_BINDINGS 16
#define MAX_PREDS 32
#define MAX_VIEWS 16

typedef struct {
    int  nrel, rel[MAX_BINDINGS];
    int  nfilter;
    filter_pred filters[MAX_PREDS];
    int  njoin;
    join_pred joins[MAX_PREDS];
    int  nview, view_b[MAX_VIEWS], view_c[MAX_VIEWS];
    char line[MAX_VIEWS * 21 + 8];      /* result: up to MAX_VIEWS 20-digit sums */
} query_t;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static int g_timing = 0;   /* B200_TIMING=1: phase times on stderr */

static relation_map *g_map   = NULL;
static int           g_nrel  = 0;
static int           g_gpus  = 1;
/* -g N: g_shard[r][j][g] = DEVICE pointer of rows [first(g), first(g + 1)) of column j of relation r on GPU g */
static uint64_t   ****g_shard = NULL;

static uint64_t shard_first(uint64_t n, int g) { return g >= g_gpus ? n : (n / (uint64_t)g_gpus) * (uint64_t)g; }

static int shard_relations(void) {
    g_shard = calloc((size_t)g_nrel, sizeof *g_shard);
    for (int r = 0; r < g_nrel; ++r) {
        g_shard[r] = calloc(g_map[r].num_columns, sizeof **g_shard);
        for (uint64_t j = 0; j < g_map[r].num_columns; ++j) g_shard[r][j] = calloc((size_t)g_gpus, sizeof(uint64_t *));
    }
    for (int g = 0; g < g_gpus; ++g) {
        if (b200_set_thread_device(g)) return 1;
        for (int r = 0; r < g_nrel; ++r) {
            const uint64_t a = shard_first(g_map[r].num_tuples, g), b = shard_first(g_map[r].num_tuples, g + 1);
            for (uint64_t j = 0; j < g_map[r].num_columns; ++j) {
                uint64_t *d = b200_device_malloc((b - a) * 8);
                if (!d || b200_copy_to_device(d, g_map[r].columns[j] + a, (b - a) * 8)) return 1;
                g_shard[r][j][g] = d;
            }
        }
    }
    return b200_set_thread_device(-1);
}

/* config 2's shape: one join of two different relations, no filter, at most one SUM per side, 32-bit keys and values */
static int try_multi_gpu(query_t *q) {
    if (g_gpus < 2 || q->nrel != 2 || q->njoin != 1 || q->nfilter != 0 || q->rel[0] == q->rel[1]) return 0;
    join_pred *j = &q->joins[0];
    if (j->relation1 == j->relation2) return 0;
    int kcol[2], vcol[2] = {-1, -1};
    kcol[j->relation1] = j->column1;
    kcol[j->relation2] = j->column2;
    for (int i = 0; i < q->nview; ++i) {
        if (vcol[q->view_b[i]] >= 0) return 0;                       /* two SUMs on one side: single-GPU path */
        vcol[q->view_b[i]] = q->view_c[i];
    }
    for (int b = 0; b < 2; ++b) {
        relation_map *rm = &g_map[q->rel[b]];
        if (rm->col_stats[kcol[b]].u >= 0xFFFFFFFFull) return 0;
        if (vcol[b] >= 0 && rm->col_stats[vcol[b]].u > 0xFFFFFFFFull) return 0;
        if (rm->num_tuples < (uint64_t)g_gpus) return 0;
    }
    const int bb = g_map[q->rel[0]].num_tuples <= g_map[q->rel[1]].num_tuples ? 0 : 1, pb = 1 - bb;   /* build = smaller */
    const uint64_t *bk[8], *bs[8], *pk[8], *ps[8];
    uint64_t        nb[8], np[8];
    for (int g = 0; g < g_gpus; ++g) {
        const int rb = q->rel[bb], rp = q->rel[pb];
        bk[g] = g_shard[rb][kcol[bb]][g];
        pk[g] = g_shard[rp][kcol[pb]][g];
        bs[g] = vcol[bb] >= 0 ? g_shard[rb][vcol[bb]][g] : NULL;
        ps[g] = vcol[pb] >= 0 ? g_shard[rp][vcol[pb]][g] : NULL;
        nb[g] = shard_first(g_map[rb].num_tuples, g + 1) - shard_first(g_map[rb].num_tuples, g);
        np[g] = shard_first(g_map[rp].num_tuples, g + 1) - shard_first(g_map[rp].num_tuples, g);
    }
    uint64_t sums[2] = {0, 0}, matches = 0;
    double   ms      = 0.0;
    const int steps  = g_timing ? 20 : 0;
    if (b200_join_sum_multi(g_gpus, B200_PLAN_BROADCAST, bk, bs, nb, pk, ps, np, steps, sums, &matches, &ms)) {
        fprintf(stderr, "b200_engine: multi-GPU plan failed (%s): single-GPU path\n", b200_last_error());
        return 0;
    }
    if (g_timing)
        fprintf(stderr, "b200_engine: multi-GPU join on %d GPUs, %lu matches, %.4f ms per step (%d steps)\n", g_gpus,
                (unsigned long)matches, ms, steps);
    /* sums come back as {SUM(build column)?, SUM(probe column)?}; print them in the query's view order */
    uint64_t by_binding[2] = {0, 0};
    int      k = 0;
    if (vcol[bb] >= 0) by_binding[bb] = sums[k++];
    if (vcol[pb] >= 0) by_binding[pb] = sums[k];
    char *p = q->line;
    for (int i = 0; i < q->nview; ++i) p += sprintf(p, "%s%lu", i ? " " : "", (unsigned long)by_binding[q->view_b[i]]);
    return 1;
}

/* ---- relation loading (relation_map.c:13-88) ------------------------------------------------------ */
static int load_relation(const char *path, relation_map *rm) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) { perror(path); return 1; }
    struct stat sb;
    if (fstat(fd, &sb) < 0) { perror("fstat"); return 1; }
    uint64_t *base = mmap(NULL, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (base == MAP_FAILED) { perror("mmap"); return 1; }
    rm->num_tuples  = base[0];
    rm->num_columns = base[1];
    rm->columns     = malloc(rm->num_columns * sizeof(uint64_t *));
    rm->col_stats   = calloc(rm->num_columns, sizeof(column_stats));
    for (uint64_t j = 0; j < rm->num_columns; ++j) rm->columns[j] = base + 2 + j * rm->num_tuples;
    return 0;
}

/* ---- query parsing (query.c:44-249) ----------------------------------------------------------------- */
static int parse_query(const char *text, query_t *q) {
    char buf[1024];
    strncpy(buf, text, sizeof buf - 1);
    buf[sizeof buf - 1] = 0;
    char *save = NULL;
    char *rels = strtok_r(buf, "|", &save), *preds = strtok_r(NULL, "|", &save), *views = strtok_r(NULL, "|\n", &save);
    if (!rels || !preds || !views) return 1;
    memset(q, 0, sizeof *q);
    char *s2 = NULL;
    for (char *t = strtok_r(rels, " ", &s2); t; t = strtok_r(NULL, " ", &s2)) {
        if (q->nrel == MAX_BINDINGS) return 1;
        const int r = atoi(t);
        if (r < 0 || r >= g_nrel) return 1;
        q->rel[q->nrel++] = r;
    }
    for (char *p = strtok_r(preds, "&", &s2); p; p = strtok_r(NULL, "&", &s2)) {
        int b1, c1, b2, c2, k;
        char op;
        if (sscanf(p, "%d.%d=%d.%d", &b1, &c1, &b2, &c2) == 4) {
            if (q->njoin == MAX_PREDS || b1 < 0 || b1 >= q->nrel || b2 < 0 || b2 >= q->nrel) return 1;
            if (c1 < 0 || (uint64_t)c1 >= g_map[q->rel[b1]].num_columns || c2 < 0 ||
                (uint64_t)c2 >= g_map[q->rel[b2]].num_columns)
                return 1;
            q->joins[q->njoin++] = (join_pred){b1, b2, c1, c2};          /* tail of the list: textual order */
        } else if (sscanf(p, "%d.%d%c%d", &b1, &c1, &op, &k) == 4 && (op == '<' || op == '>' || op == '=')) {
            if (q->nfilter == MAX_PREDS || b1 < 0 || b1 >= q->nrel) return 1;
            if (c1 < 0 || (uint64_t)c1 >= g_map[q->rel[b1]].num_columns) return 1;
            /* filters go to the list head => they run in reverse textual order (query.c:150-157) */
            memmove(q->filters + 1, q->filters, (size_t)q->nfilter * sizeof(filter_pred));
            q->filters[0] = (filter_pred){b1, c1, k, op};
            q->nfilter++;
        } else {
            return 1;
        }
    }
    for (char *v = strtok_r(views, " \n", &s2); v; v = strtok_r(NULL, " \n", &s2)) {
        if (q->nview == MAX_VIEWS || strlen(v) != 3 || v[1] != '.') return 1;
        const int b = v[0] - '0', c = v[2] - '0';  /* single digits, inter_res.c:325-327 */
        if (b < 0 || b >= q->nrel || c < 0 || (uint64_t)c >= g_map[q->rel[b]].num_columns) return 1;
        q->view_b[q->nview] = b;
        q->view_c[q->nview] = c;
        q->nview++;
    }
    return q->nrel == 0 || q->nview == 0 || q->njoin + q->nfilter == 0;
}

static void null_line(query_t *q) {
    char *p = q->line;
    for (int i = 0; i < q->nview; ++i) p += sprintf(p, "%sNULL", i ? " " : "");
}

/* ---- ExecuteQuery (query.c:325-467) over the operator API ------------------------------------------- */
static void execute_query(query_t *q) {
    if (try_multi_gpu(q)) return;
    inter_res *inter = NULL;
    InitInterResults(&inter, q->nrel);
    for (int i = 0; i < q->nfilter; ++i) {
        result *res = Filter(inter, &q->filters[i], g_map, q->rel);
        if (!res) { null_line(q); FreeInterResults(inter); return; }           /* query.c:360-369 */
        InsertSingleRowIdsToInterResult(&inter, q->filters[i].relation, res);
        FreeResult(res);
    }
    /* joins: a predicate whose two bindings already share a node is a filter on the intermediate */
    int done[MAX_PREDS] = {0}, left = q->njoin;
    while (left) {
        int pick = -1;
        /* prefer a predicate connected to what has been joined/filtered so far (keeps intermediates small) */
        for (int i = 0; i < q->njoin && pick < 0; ++i) {
            if (done[i]) continue;
            for (inter_res *n = inter; n; n = n->next)
                if (n->data->num_tuples && (n->data->table[q->joins[i].relation1] || n->data->table[q->joins[i].relation2]))
                    pick = i;
        }
        for (int i = 0; i < q->njoin && pick < 0; ++i)
            if (!done[i]) pick = i;
        join_pred *j = &q->joins[pick];
        done[pick]   = 1;
        --left;
        if (j->relation1 == j->relation2) {
            result *res = SelfJoin(j->relation1, j->column1, j->column2, &inter, g_map, q->rel);
            if (!res) { null_line(q); FreeInterResults(inter); return; }
            InsertSingleRowIdsToInterResult(&inter, j->relation1, res);
            FreeResult(res);
            continue;
        }
        if (AreActiveInInter(inter, j->relation1, j->relation2)) {
            JoinInterNode(&inter, g_map, j->relation1, j->column1, j->relation2, j->column2, q->rel);
            continue;
        }
        relation *r1 = GetRelation(j->relation1, j->column1, inter, g_map, q->rel);
        relation *r2 = GetRelation(j->relation2, j->column2, inter, g_map, q->rel);
        result   *res = RadixHashJoin(r1, r2, NULL);
        FreeRelation(r1);
        FreeRelation(r2);
        if (!res) { null_line(q); FreeInterResults(inter); return; }           /* query.c:439-449 */
        InsertJoinToInterResults(inter, j->relation1, j->relation2, res);
        FreeResult(res);
        if (inter->next) MergeInterNodes(&inter);
    }
    if (inter->next) CartesianInterResults(&inter);
    /* CalculateQueryResults without the printf: lines are printed in order when the batch ends */
    char               v0[MAX_VIEWS][4];
    char              *vp[MAX_VIEWS];
    query_string_array views = {vp, q->nview};
    for (int i = 0; i < q->nview; ++i) {
        snprintf(v0[i], sizeof v0[i], "%d.%d", q->view_b[i], q->view_c[i]);
        vp[i] = v0[i];
    }
    batch_listnode node = {q->nrel, q->rel, NULL, &views, NULL};
    uint64_t       sums[MAX_VIEWS], rows = 0;
    b200_calculate_sums(inter, g_map, &node, sums, &rows);
    if (b200_last_result_null()) {                     /* a filter fused into a join let nothing through */
        null_line(q);
        FreeInterResults(inter);
        return;
    }
    char *p = q->line;
    for (int i = 0; i < q->nview; ++i) p += sprintf(p, "%s%lu", i ? " " : "", (unsigned long)sums[i]);
    FreeInterResults(inter);
}

/* ---- the scheduler: jobs = whole queries, workers = threads that own a CUDA stream -------------------
 * scheduler.c:9-132 keeps a pool of pthreads alive for the whole process (SchedulerInit once, handler.c:61-63)
 * and feeds it jobs; so does this one: the workers are created once, every batch is published under the
 * mutex, and the main thread waits until the batch's last query has been answered (Barrier, scheduler.c:76-86).
 * A worker keeps its library context (CUDA stream, pinned scratch) across batches. */
typedef struct {
    pthread_mutex_t mu;
    pthread_cond_t  work, done;
    query_t        *queries;
    int             n, next, finished, stop;
    int             nworkers;
    pthread_t       th[64];
} pool_t;
static pool_t g_pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER,
                        NULL, 0, 0, 0, 0, 0, {0}};

static void *worker(void *arg) {
    pool_t *p = arg;
    pthread_mutex_lock(&p->mu);
    for (;;) {
        while (!p->stop && p->next >= p->n) pthread_cond_wait(&p->work, &p->mu);
        if (p->stop) break;
        const int i = p->next++;
        pthread_mutex_unlock(&p->mu);
        execute_query(&p->queries[i]);
        pthread_mutex_lock(&p->mu);
        if (++p->finished == p->n) pthread_cond_signal(&p->done);
    }
    pthread_mutex_unlock(&p->mu);
    return NULL;
}

static void pool_start(int workers) {
    g_pool.nworkers = workers;
    for (int w = 0; w < workers; ++w) pthread_create(&g_pool.th[w], NULL, worker, &g_pool);
}

static void pool_stop(void) {
    pthread_mutex_lock(&g_pool.mu);
    g_pool.stop = 1;
    pthread_cond_broadcast(&g_pool.work);
    pthread_mutex_unlock(&g_pool.mu);
    for (int w = 0; w < g_pool.nworkers; ++w) pthread_join(g_pool.th[w], NULL);
}

/* B200_TIMING=2: every query's time and the library's kernel timers on stderr (one worker: the timers are per stream) */
static void execute_query_timed(query_t *q, int i) {
    static const char *names[] = {"filter", "hist_b", "hist_p", "scatter_b", "scatter_p", "scatter_pc", "filter_fused",
                                  "join", "join_write", "overflow"};
    b200_set_profiling(1);               /* forgets the timers of the previous query */
    const double t0 = now_s();
    execute_query(q);
    char  buf[512];
    char *p = buf;
    for (unsigned k = 0; k < sizeof names / sizeof names[0]; ++k) {
        int          scopes = 0;
        const double ms     = b200_sum_kernel_ms(names[k], &scopes);
        if (ms >= 0) p += sprintf(p, " %s=%.3f(x%d)", names[k], ms, scopes);
    }
    fprintf(stderr, "b200_engine: query %d: %.3f ms |%s | %s\n", i, (now_s() - t0) * 1e3, buf, q->line);
}

static void run_batch(query_t *queries, int n) {
    if (g_pool.nworkers <= 1) {
        for (int i = 0; i < n; ++i) {
            if (g_timing > 1) execute_query_timed(&queries[i], i);
            else execute_query(&queries[i]);
        }
    } else if (n > 0) {
        pthread_mutex_lock(&g_pool.mu);
        g_pool.queries  = queries;
        g_pool.n        = n;
        g_pool.next     = 0;
        g_pool.finished = 0;
        pthread_cond_broadcast(&g_pool.work);
        while (g_pool.finished < n) pthread_cond_wait(&g_pool.done, &g_pool.mu);
        g_pool.n = g_pool.next = 0;
        pthread_mutex_unlock(&g_pool.mu);
    }
    for (int i = 0; i < n; ++i) puts(queries[i].line);      /* submission order, like the reference */
    fflush(stdout);
}

int main(int argc, char **argv) {
    int workers = 4;
    if (getenv("B200_WORKERS")) workers = atoi(getenv("B200_WORKERS"));
    if (getenv("B200_GPUS")) g_gpus = atoi(getenv("B200_GPUS"));
    for (int a = 1; a + 1 < argc; a += 2) {
        if (!strcmp(argv[a], "-w")) workers = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "-g")) g_gpus = atoi(argv[a + 1]);
    }
    if (workers < 1) workers = 1;
    if (workers > 64) workers = 64;
    if (g_gpus < 1) g_gpus = 1;
    if (g_gpus > 8) g_gpus = 8;

    char buff[1024];
    int  cap = 16;
    g_map    = malloc((size_t)cap * sizeof(relation_map));
    while (scanf("%1023s", buff) == 1 && strcmp(buff, "Done")) {            /* handler.c:27-48 */
        if (g_nrel == cap) g_map = realloc(g_map, (size_t)(cap *= 2) * sizeof(relation_map));
        if (load_relation(buff, &g_map[g_nrel])) return 1;
        ++g_nrel;
    }
    g_timing  = getenv("B200_TIMING") ? (atoi(getenv("B200_TIMING")) > 1 ? 2 : 1) : 0;
    double   t0 = now_s();
    uint64_t reserved = 0;
    b200_init(-1);
    double t1 = now_s();
    b200_register_relations(g_map, g_nrel);     /* the untimed preparation phase: columns go to HBM once */
    b200_compute_column_stats(g_map, g_nrel);   /* relation_map.c:53-83's min / max / distinct, computed on the GPU */
    {   /* still the preparation phase: room for the intermediates (eight times the relations, at least 8 GB; $B200_RESERVE_GB) reserved in the pool */
        uint64_t total = 0;
        for (int r = 0; r < g_nrel; ++r) total += 8ull * g_map[r].num_tuples * g_map[r].num_columns;
        const char *gb = getenv("B200_RESERVE_GB");
        uint64_t want = gb ? (uint64_t)atoll(gb) << 30 : (8 * total > (8ull << 30) ? 8 * total : 8ull << 30);
        reserved = b200_reserve_device_memory(want);
    }
    if (g_gpus > 1) {
        if (shard_relations()) { fprintf(stderr, "b200_engine: cannot shard over %d GPUs: %s\n", g_gpus, b200_last_error()); return 1; }
        workers = 1;                            /* the multi-GPU plan runs one query at a time over all GPUs */
    }
    if (g_timing > 1) b200_set_profiling(1);
    if (g_timing) fprintf(stderr, "b200_engine: CUDA start-up %.3f s, upload of %d relations, statistics and %.1f GB reserved %.3f s\n", t1 - t0, g_nrel, (double)reserved / 1073741824.0, now_s() - t1);

    if (workers > 1) pool_start(workers);
    query_t *batch = NULL;
    int      nq = 0, qcap = 0;
    while (fgets(buff, sizeof buff, stdin)) {                              /* handler.c:66-96 */
        if (strlen(buff) < 2) continue;
        if (!strcmp(buff, "Exit\n")) break;
        if (!strcmp(buff, "F\n")) {
            double tb = now_s();
            run_batch(batch, nq);
            if (g_timing) fprintf(stderr, "b200_engine: batch of %d queries on %d workers: %.3f s\n", nq, workers, now_s() - tb);
            nq = 0;
            continue;
        }
        if (nq == qcap) batch = realloc(batch, (size_t)(qcap = qcap ? 2 * qcap : 64) * sizeof(query_t));
        if (parse_query(buff, &batch[nq])) { fprintf(stderr, "cannot parse query: %s", buff); return 2; }
        ++nq;
    }
    if (nq) run_batch(batch, nq);
    pool_stop();
    b200_shutdown();
    return 0;
}
