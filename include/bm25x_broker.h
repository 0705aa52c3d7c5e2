/* bm25x_broker.h — batching broker in front of bm25x_search_batch.
 *
 * The reference answers ONE query per index scan: `DefaultBuilder::build` calls `bm25::search(&index, limit, &vector,
 * filter)` from the backend process that runs the scan (src/index/bm25/scanners/default.rs:93-134, driven by the AM
 * callbacks amrescan / amgettuple, src/index/bm25/am/mod.rs:345-433).  The GPU path earns its keep on batches of 10^3..10^5
 * queries, and one CUDA context per Postgres backend is not viable — so the integration runs ONE process (background
 * worker / sidecar) that owns the CUDA context and the bm25x_index handles, and every backend hands its
 * (tokens, limit) to it.  This is that process's core: concurrent callers enqueue single queries into a bounded request
 * ring; one worker thread drains the ring, coalesces what it finds into a batch (per limit class), runs
 * bm25x_search_batch once per batch and scatters the rows back to the callers.
 *
 * What is here: the in-process ring + worker + scatter, with the search behind a function pointer (so the batching logic
 * is testable without a GPU).  What a deployment adds around it: the shared-memory transport between the backends and
 * this process (the request / response structs below are plain data for that reason) and the pgrx glue.
 *
 * Results are exactly those of bm25x_search_batch: a request with limit k served from a batch run at a larger k'
 * receives the first k rows of its k' rows — identical, because results are totally ordered (score desc, doc id asc). */
#ifndef BM25X_BROKER_H
#define BM25X_BROKER_H

#include <stdint.h>

#include "bm25x.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bm25x_broker bm25x_broker;

/* The call the broker batches (signature of bm25x_search_batch minus handle / prefilter / statistics). */
typedef int (*bm25x_broker_backend)(void *ctx, uint32_t nq, const uint32_t *q_off, const uint32_t *q_terms, uint32_t k,
                                    uint32_t *out_doc, float *out_score, double *out_score64, uint16_t *out_payload,
                                    uint32_t *out_n);

typedef struct {
    uint32_t max_batch;   /* queries per backend call at most (0: 65536) */
    uint32_t max_wait_us; /* how long the worker waits for more requests after the first one of a batch (0: 200) */
    uint32_t ring_slots;  /* capacity of the request ring; callers block while it is full (0: 2 x max_batch) */
    uint32_t reserved;
} bm25x_broker_options;

typedef struct {
    uint64_t requests;        /* queries answered */
    uint64_t batches;         /* backend calls */
    uint64_t max_batch_seen;  /* largest batch so far */
    uint64_t ring_full_waits; /* times a caller found the ring full */
    uint64_t rejected;        /* requests refused before batching (limit 0 / too large, too many tokens) */
} bm25x_broker_stats;

/* Broker over an index handle: the backend is bm25x_search_batch(idx, ...).  The worker thread makes the CUDA calls. */
int bm25x_broker_create(bm25x_index *idx, const bm25x_broker_options *opt, bm25x_broker **out);
/* Broker over any backend (tests, other engines). */
int bm25x_broker_create_with_backend(bm25x_broker_backend fn, void *ctx, const bm25x_broker_options *opt,
                                     bm25x_broker **out);
/* One query, the shape of bm25::search(&index, limit, &vector, filter) (search.rs:28-35): term ordinals in (unknown /
 * duplicate ones are handled by the backend as in bm25x_search_batch), at most `limit` rows out, best first.
 * Blocks until the batch holding the request has been answered.  Thread-safe.  out_score64 / out_payload may be NULL.
 * Status codes as bm25x_search_batch; BM25X_ERR_LIMIT_ZERO for limit 0 (scanners/default.rs:114-116). */
int bm25x_broker_search(bm25x_broker *b, const uint32_t *terms, uint32_t n_terms, uint32_t limit, uint32_t *out_doc,
                        double *out_score64, uint16_t *out_payload, uint32_t *out_n);
int bm25x_broker_get_stats(const bm25x_broker *b, bm25x_broker_stats *out);
/* Answers what is queued, then stops the worker. */
void bm25x_broker_destroy(bm25x_broker *b);

#ifdef __cplusplus
}
#endif
#endif
