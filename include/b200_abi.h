/*
 * b200_abi.h — data types that cross the operator boundary.
 *
 * The join hot path of VagelisN/Sigmod-2018 is called from exactly one place,
 * query.c:ExecuteQuery (reference query.c:325-467).  The structs below are
 * the ones that caller passes to / receives from the operators.  Their field
 * order and widths are the reference's (structs.h, cited per type) because the
 * reference's own compiled query.o / handler.o must be able to link against
 * this library unchanged.  Everything an operator returns
 * (`relation`, `result`, `inter_res->data`) is opaque to that caller — it only
 * passes the pointers back in and frees them through FreeResult /
 * FreeRelation / FreeInterResults — so the B200 library keeps DEVICE pointers
 * and private bookkeeping behind those public prefixes.
 *
 * All values on the path are unsigned 64-bit integers (keys, payloads, sums
 * mod 2^64).  There is no floating point on the hot path.
 */
#ifndef B200_ABI_H
#define B200_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference structs.h:15-19 — AoS join tuple (the reference materialises
 * these on the host; this library never does, see DESIGN.md "K2 fused away") */
typedef struct tuple {
    uint64_t value;
    uint64_t row_id;
} tuple;

/* reference structs.h:25-29 — public prefix of what GetRelation returns. */
typedef struct relation {
    tuple   *tuples;      /* NULL in this library: the key vector stays a lazy
                             (column, row-id list) descriptor in HBM */
    uint64_t num_tuples;
} relation;

/* reference structs.h:37-43 — public prefix of what Filter / RadixHashJoin
 * return (a linked list of byte buffers on the reference side). */
typedef struct result {
    char          *buff;         /* NULL in this library */
    struct result *next;         /* always NULL in this library */
    uint64_t       current_load; /* number of row ids / pairs held */
} result;

/* reference structs.h:46-50 */
typedef struct result_tuple {
    uint64_t row_idR;
    uint64_t row_idS;
} result_tuple;

/* reference structs.h:97-101 — one SoA table of row ids per query binding;
 * table[b] == NULL means binding b is not part of this node yet. */
typedef struct inter_data {
    uint64_t   num_tuples;
    uint64_t **table;   /* in this library: DEVICE pointers to 32-bit row ids,
                           stored behind the reference's pointer type */
} inter_data;

/* reference structs.h:106-111 — the caller reads only ->next
 * (query.c:453, query.c:462). */
typedef struct intermediate_result {
    struct inter_data          *data;
    int                         num_of_relations;
    struct intermediate_result *next;
} inter_res;

/* reference structs.h:121-127 */
typedef struct column_stats {
    uint64_t l;   /* min */
    uint64_t u;   /* max */
    double   f;   /* row count */
    double   d;   /* distinct count */
} column_stats;

/* reference structs.h:133-139 — host view of one loaded relation; `columns`
 * point into the mmap'd file (relation_map.c:46-50).  The library keys its
 * device copies by these host column pointers (b200_register_relations). */
typedef struct relation_map {
    uint64_t      num_tuples;
    uint64_t      num_columns;
    uint64_t    **columns;
    column_stats *col_stats;
} relation_map;

/* reference structs.h:142-148 — constant is a 32-bit int (query.c:239). */
typedef struct filter_pred {
    int  relation;
    int  column;
    int  value;
    char comperator;   /* '<', '>' or '=' (spelling is the reference's) */
} filter_pred;

/* reference structs.h:152-158 */
typedef struct join_pred {
    int relation1;
    int relation2;
    int column1;
    int column2;
} join_pred;

/* reference structs.h:166-171 */
typedef struct predicates_listnode {
    filter_pred                *filter_p;
    join_pred                  *join_p;
    struct predicates_listnode *next;
} predicates_listnode;

/* reference structs.h:177-181 — projection strings "b.c" */
typedef struct query_string_array {
    char **data;
    int    num_of_elements;
} query_string_array;

/* reference structs.h:188-195 */
typedef struct query_batch_listnode {
    int                          num_of_relations;
    int                         *relations;
    predicates_listnode         *predicate_list;
    query_string_array          *views;
    struct query_batch_listnode *next;
} batch_listnode;

/* reference structs.h:211-225.  RadixHashJoin receives it (rhjoin.h:9) only
 * to fan work out over pthreads; the B200 library launches kernels on the
 * calling thread's CUDA stream instead and never dereferences it. */
struct scheduler;
typedef struct scheduler scheduler;

#ifdef __cplusplus
}
#endif
#endif /* B200_ABI_H */
