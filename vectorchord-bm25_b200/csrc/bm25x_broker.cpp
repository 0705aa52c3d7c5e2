// bm25x_broker.cpp — batching broker in front of bm25x_search_batch (include/bm25x_broker.h).
//
// Replaces, for a deployment, the one-query-per-scan call of the reference (DefaultBuilder::build →
// bm25::search, src/index/bm25/scanners/default.rs:117-129): callers enqueue single queries into a bounded ring, ONE
// worker thread (the owner of the CUDA context) coalesces them into batches per limit class and scatters the rows back.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/bm25x_broker.h"

void bm25x_set_error(const char *fmt, ...);

namespace {

// One pending query: lives on its caller's stack for the duration of bm25x_broker_search.
struct Request {
    const uint32_t *terms;
    uint32_t n_terms, limit;
    uint32_t *out_doc;
    double *out_score64;
    uint16_t *out_payload;
    uint32_t *out_n;
    int rc = BM25X_OK;
    bool done = false;
};

// Limit classes: the requests of one backend call run at the largest limit of their class and each takes the first
// `limit` rows of its result (results are totally ordered: a prefix of a longer top-k IS the shorter top-k).  The classes
// follow the pool capacities of the search kernel (bm25x_search.cu: k <= 32, <= 128 seeded, <= 224, <= 1024, beyond).
int limit_class(uint32_t k) { return k <= 32 ? 0 : k <= 128 ? 1 : k <= 224 ? 2 : k <= 1024 ? 3 : 4; }
constexpr int kLimitClasses = 5;

int index_backend(void *ctx, uint32_t nq, const uint32_t *q_off, const uint32_t *q_terms, uint32_t k, uint32_t *out_doc,
                  float *out_score, double *out_score64, uint16_t *out_payload, uint32_t *out_n) {
    return bm25x_search_batch((bm25x_index *)ctx, nq, q_off, q_terms, k, nullptr, out_doc, out_score, out_score64, out_payload,
                              out_n, nullptr);
}

}  // namespace

struct bm25x_broker {
    bm25x_broker_backend fn = nullptr;
    void *ctx = nullptr;
    uint32_t max_batch = 65536, max_wait_us = 200, ring_slots = 131072;
    // the request ring (bounded; callers block while it is full)
    std::vector<Request *> ring;
    size_t head = 0, count = 0;
    mutable std::mutex mu;
    std::condition_variable cv_work, cv_space, cv_done, cv_idle;
    bool stop = false;
    uint32_t inflight = 0;  // callers inside bm25x_broker_search (destroy waits for them to leave)
    bm25x_broker_stats st{};
    std::thread worker;

    void run();
    void serve(std::vector<Request *> &batch, uint32_t k);
    // batch buffers of the worker (grow-only)
    std::vector<uint32_t> q_off, q_terms, o_doc, o_n;
    std::vector<float> o_score;
    std::vector<double> o_score64;
    std::vector<uint16_t> o_payload;
};

void bm25x_broker::serve(std::vector<Request *> &batch, uint32_t k) {
    const uint32_t nq = (uint32_t)batch.size();
    q_off.resize((size_t)nq + 1);
    q_terms.clear();
    q_off[0] = 0;
    bool want_payload = false;
    for (uint32_t i = 0; i < nq; ++i) {
        q_terms.insert(q_terms.end(), batch[i]->terms, batch[i]->terms + batch[i]->n_terms);
        q_off[i + 1] = (uint32_t)q_terms.size();
        want_payload = want_payload || batch[i]->out_payload != nullptr;
    }
    const size_t slots = (size_t)nq * k;
    o_doc.resize(slots);
    o_score.resize(slots);
    o_score64.resize(slots);
    if (want_payload) o_payload.resize(slots * 3);
    o_n.resize(nq);
    if (q_terms.empty()) q_terms.push_back(0);  // (never read: every query is empty)
    const int rc = fn(ctx, nq, q_off.data(), q_terms.data(), k, o_doc.data(), o_score.data(), o_score64.data(),
                      want_payload ? o_payload.data() : nullptr, o_n.data());
    for (uint32_t i = 0; i < nq; ++i) {
        Request *r = batch[i];
        r->rc = rc;
        if (rc == BM25X_OK) {
            const uint32_t n = std::min(o_n[i], r->limit);  // the first `limit` rows of the class-wide top-k
            const size_t o = (size_t)i * k;
            if (r->out_doc) memcpy(r->out_doc, o_doc.data() + o, sizeof(uint32_t) * n);
            if (r->out_score64) memcpy(r->out_score64, o_score64.data() + o, sizeof(double) * n);
            if (r->out_payload) memcpy(r->out_payload, o_payload.data() + 3 * o, sizeof(uint16_t) * 3 * n);
            if (r->out_n) *r->out_n = n;
        }
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        for (Request *r : batch) r->done = true;
        st.requests += nq;
        st.batches += 1;
        st.max_batch_seen = std::max<uint64_t>(st.max_batch_seen, nq);
    }
    cv_done.notify_all();
}

void bm25x_broker::run() {
    using clk = std::chrono::steady_clock;
    std::vector<Request *> taken, group[kLimitClasses];
    for (;;) {
        taken.clear();
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_work.wait(lk, [&] { return count > 0 || stop; });
            if (count == 0 && stop) return;
            // the first request of a batch is here: give the others max_wait_us to join it (or fill the batch)
            const auto deadline = clk::now() + std::chrono::microseconds(max_wait_us);
            while (count < max_batch && !stop) {
                if (cv_work.wait_until(lk, deadline) == std::cv_status::timeout) break;
            }
            const size_t n = std::min<size_t>(count, max_batch);
            for (size_t i = 0; i < n; ++i) taken.push_back(ring[(head + i) % ring.size()]);
            head = (head + n) % ring.size();
            count -= n;
        }
        cv_space.notify_all();
        for (auto &g : group) g.clear();
        for (Request *r : taken) group[limit_class(r->limit)].push_back(r);
        for (auto &g : group) {
            if (g.empty()) continue;
            uint32_t k = 0;
            for (Request *r : g) k = std::max(k, r->limit);
            serve(g, k);
        }
    }
}

static int broker_make(bm25x_broker_backend fn, void *ctx, const bm25x_broker_options *opt, bm25x_broker **out) {
    if (!fn || !out) {
        bm25x_set_error("bm25x_broker_create: null argument");
        return BM25X_ERR_INVALID;
    }
    bm25x_broker *b = new bm25x_broker();
    b->fn = fn;
    b->ctx = ctx;
    if (opt) {
        if (opt->max_batch) b->max_batch = opt->max_batch;
        if (opt->max_wait_us) b->max_wait_us = opt->max_wait_us;
        b->ring_slots = opt->ring_slots ? opt->ring_slots : 2 * b->max_batch;
    }
    if (b->ring_slots < 1) b->ring_slots = 1;
    b->ring.assign(b->ring_slots, nullptr);
    b->worker = std::thread([b] { b->run(); });
    *out = b;
    return BM25X_OK;
}

extern "C" int bm25x_broker_create(bm25x_index *idx, const bm25x_broker_options *opt, bm25x_broker **out) {
    if (!idx) {
        bm25x_set_error("bm25x_broker_create: null index");
        return BM25X_ERR_INVALID;
    }
    return broker_make(index_backend, idx, opt, out);
}

extern "C" int bm25x_broker_create_with_backend(bm25x_broker_backend fn, void *ctx, const bm25x_broker_options *opt,
                                                bm25x_broker **out) {
    return broker_make(fn, ctx, opt, out);
}

extern "C" int bm25x_broker_search(bm25x_broker *b, const uint32_t *terms, uint32_t n_terms, uint32_t limit,
                                   uint32_t *out_doc, double *out_score64, uint16_t *out_payload, uint32_t *out_n) {
    if (!b || (n_terms && !terms) || !out_doc || !out_n) {
        bm25x_set_error("bm25x_broker_search: null argument");
        return BM25X_ERR_INVALID;
    }
    // what would fail the whole batch is refused here, for this caller alone
    int reject = BM25X_OK;
    if (limit == 0) {
        bm25x_set_error("number of needed rows is set to 0");  // scanners/default.rs:114-116
        reject = BM25X_ERR_LIMIT_ZERO;
    } else if (limit > BM25X_MAX_K) {
        bm25x_set_error("bm25x_broker_search: limit=%u > BM25X_MAX_K=%d", limit, BM25X_MAX_K);
        reject = BM25X_ERR_UNSUPPORTED;
    } else if (n_terms > BM25X_MAX_QUERY_TERMS) {
        bm25x_set_error("bm25x_broker_search: %u tokens > %d", n_terms, BM25X_MAX_QUERY_TERMS);
        reject = BM25X_ERR_UNSUPPORTED;
    }
    Request r;
    r.terms = terms;
    r.n_terms = n_terms;
    r.limit = limit;
    r.out_doc = out_doc;
    r.out_score64 = out_score64;
    r.out_payload = out_payload;
    r.out_n = out_n;
    std::unique_lock<std::mutex> lk(b->mu);
    if (reject != BM25X_OK) {
        b->st.rejected++;
        return reject;
    }
    b->inflight++;
    auto leave = [&](int rc) {  // (still under the lock: destroy deletes the broker only after the last caller has left)
        if (--b->inflight == 0) b->cv_idle.notify_all();
        return rc;
    };
    if (b->stop) {
        bm25x_set_error("bm25x_broker_search: the broker is shutting down");
        return leave(BM25X_ERR_INVALID);
    }
    if (b->count == b->ring.size()) {
        b->st.ring_full_waits++;
        b->cv_space.wait(lk, [&] { return b->count < b->ring.size() || b->stop; });
        if (b->stop) {
            bm25x_set_error("bm25x_broker_search: the broker is shutting down");
            return leave(BM25X_ERR_INVALID);
        }
    }
    b->ring[(b->head + b->count) % b->ring.size()] = &r;
    b->count++;
    b->cv_work.notify_one();
    b->cv_done.wait(lk, [&] { return r.done; });
    if (r.rc != BM25X_OK) bm25x_set_error("bm25x_broker_search: the batch holding this query failed with status %d", r.rc);
    return leave(r.rc);
}

extern "C" int bm25x_broker_get_stats(const bm25x_broker *b, bm25x_broker_stats *out) {
    if (!b || !out) {
        bm25x_set_error("bm25x_broker_get_stats: null argument");
        return BM25X_ERR_INVALID;
    }
    std::lock_guard<std::mutex> lk(b->mu);
    *out = b->st;
    return BM25X_OK;
}

extern "C" void bm25x_broker_destroy(bm25x_broker *b) {
    if (!b) return;
    {
        std::lock_guard<std::mutex> lk(b->mu);
        b->stop = true;  // the worker answers what is queued, then returns; new callers are refused
    }
    b->cv_work.notify_all();
    b->cv_space.notify_all();
    if (b->worker.joinable()) b->worker.join();
    {   // callers that were being answered are still on their way out
        std::unique_lock<std::mutex> lk(b->mu);
        b->cv_idle.wait(lk, [&] { return b->inflight == 0; });
    }
    delete b;
}
