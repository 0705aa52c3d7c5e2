// bm25x_synth.cpp — synthetic corpus / query generator (bench + test utility, host side, OpenMP).
//
// Spec (mirrors the reference's fuzz generator tests/fuzz:168-205: L draws with
// replacement from the vocabulary, duplicates aggregated into tf, document length = L, doc id =
// generation order).  Counter-based so every document is reproducible on its own:
//   h(seed, item)   = sm64(seed ^ sm64(item))
//   draw(seed,i,j)  = sm64(h + j * 0x9E3779B97F4A7C15)
//   uniform term    = ((u >> 32) * vocab) >> 32
//   Zipf(s) term    = first rank r with u <= floor(2^64 * CDF(r))
//   length          = len_min + ((draw(seed,i,0xFFFFFFFF) >> 32) * (len_max-len_min+1)) >> 32
// The oracle (oracle/bm25_oracle.c) restates the same spec independently; tests compare the two.
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/bm25x.h"

int bm25x_host_threads(int cap);  // bm25x_index.cu: affinity mask capped by the cgroup CPU quota

void bm25x_set_error(const char *fmt, ...);

static inline uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t draw(uint64_t seed, uint64_t item, uint32_t j) {
    uint64_t h = sm64(seed ^ sm64(item));
    return sm64(h + (uint64_t)j * 0x9E3779B97F4A7C15ull);
}
static void zipf_thresholds(uint32_t vocab, double s, std::vector<uint64_t> &thr) {
    thr.resize(vocab);
    double total = 0.0;
    for (uint32_t r = 0; r < vocab; r++) total += pow((double)(r + 1), -s);
    double acc = 0.0;
    for (uint32_t r = 0; r < vocab; r++) {
        acc += pow((double)(r + 1), -s);
        double c = acc / total;
        thr[r] = (c >= 1.0 || r + 1 == vocab) ? 0xFFFFFFFFFFFFFFFFull : (uint64_t)ldexp(c, 64);
    }
}
static inline uint32_t draw_term(uint64_t u, uint32_t vocab, const uint64_t *thr) {
    if (!thr) return (uint32_t)(((u >> 32) * (uint64_t)vocab) >> 32);
    uint32_t lo = 0, hi = vocab - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (u <= thr[mid]) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}
static inline uint32_t doc_length(uint64_t seed, uint32_t d, uint32_t len_min, uint32_t len_max) {
    if (len_max <= len_min) return len_min;
    uint64_t u = draw(seed, d, 0xFFFFFFFFu);
    return len_min + (uint32_t)(((u >> 32) * (uint64_t)(len_max - len_min + 1)) >> 32);
}
// sorted draws of one document into buf[0..L)
static inline uint32_t gen_doc(uint64_t seed, uint32_t d, uint32_t vocab, uint32_t len_min, uint32_t len_max,
                               const uint64_t *thr, uint32_t *buf) {
    uint32_t L = doc_length(seed, d, len_min, len_max);
    uint64_t h = sm64(seed ^ sm64((uint64_t)d));
    for (uint32_t j = 0; j < L; j++) buf[j] = draw_term(sm64(h + (uint64_t)j * 0x9E3779B97F4A7C15ull), vocab, thr);
    std::sort(buf, buf + L);
    return L;
}

extern "C" void bm25x_synth_free(bm25x_synth_corpus *c) {
    if (!c) return;
    free(c->doc_len);
    free(c->post_off);
    free(c->post_doc);
    free(c->post_tf);
    memset(c, 0, sizeof(*c));
}

extern "C" int bm25x_synth_generate(uint64_t seed, uint32_t n_docs, uint32_t vocab, uint32_t len_min, uint32_t len_max,
                                    double zipf_s, int nthreads, bm25x_synth_corpus *out) {
    if (!out || n_docs == 0 || vocab == 0 || len_max < len_min) {
        bm25x_set_error("bm25x_synth_generate: bad arguments");
        return BM25X_ERR_INVALID;
    }
    memset(out, 0, sizeof(*out));
    if (nthreads < 1) nthreads = bm25x_host_threads(0);
    std::vector<uint64_t> thrv;
    const uint64_t *thr = nullptr;
    if (zipf_s > 0.0) {
        zipf_thresholds(vocab, zipf_s, thrv);
        thr = thrv.data();
    }
    const int nchunks = std::max(1, std::min<int>(nthreads * 4, (int)std::min<uint32_t>(n_docs, 1024)));
    const uint32_t per = (n_docs + nchunks - 1) / nchunks;
    std::vector<uint32_t> cnt((size_t)nchunks * vocab, 0);
    out->doc_len = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n_docs);
    out->post_off = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)vocab + 1));
    if (!out->doc_len || !out->post_off) {
        bm25x_synth_free(out);
        bm25x_set_error("bm25x_synth_generate: out of host memory");
        return BM25X_ERR_OOM;
    }
    // pass 1: per-chunk document frequencies
#pragma omp parallel num_threads(nthreads)
    {
        std::vector<uint32_t> buf(len_max ? len_max : 1);
#pragma omp for schedule(dynamic, 1)
        for (int c = 0; c < nchunks; c++) {
            uint32_t d0 = (uint32_t)c * per, d1 = std::min<uint64_t>((uint64_t)d0 + per, n_docs);
            uint32_t *cc = cnt.data() + (size_t)c * vocab;
            for (uint32_t d = d0; d < d1; d++) {
                uint32_t L = gen_doc(seed, d, vocab, len_min, len_max, thr, buf.data());
                out->doc_len[d] = L;
                for (uint32_t j = 0; j < L; j++)
                    if (j == 0 || buf[j] != buf[j - 1]) cc[buf[j]]++;
            }
        }
    }
    // offsets: term-major, chunk-minor ⇒ doc ids ascend inside a term
    uint64_t acc = 0;
    std::vector<uint64_t> cur((size_t)nchunks * vocab);
    for (uint32_t t = 0; t < vocab; t++) {
        out->post_off[t] = acc;
        for (int c = 0; c < nchunks; c++) {
            cur[(size_t)c * vocab + t] = acc;
            acc += cnt[(size_t)c * vocab + t];
        }
    }
    out->post_off[vocab] = acc;
    out->n_docs = n_docs;
    out->n_terms = vocab;
    out->n_postings = acc;
    out->post_doc = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(acc ? acc : 1));
    out->post_tf = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(acc ? acc : 1));
    if (!out->post_doc || !out->post_tf) {
        bm25x_synth_free(out);
        bm25x_set_error("bm25x_synth_generate: out of host memory (%llu postings)", (unsigned long long)acc);
        return BM25X_ERR_OOM;
    }
    // pass 2: regenerate and scatter
#pragma omp parallel num_threads(nthreads)
    {
        std::vector<uint32_t> buf(len_max ? len_max : 1);
#pragma omp for schedule(dynamic, 1)
        for (int c = 0; c < nchunks; c++) {
            uint32_t d0 = (uint32_t)c * per, d1 = std::min<uint64_t>((uint64_t)d0 + per, n_docs);
            uint64_t *cc = cur.data() + (size_t)c * vocab;
            for (uint32_t d = d0; d < d1; d++) {
                uint32_t L = gen_doc(seed, d, vocab, len_min, len_max, thr, buf.data());
                uint32_t j = 0;
                while (j < L) {
                    uint32_t t = buf[j], tf = 1;
                    while (j + tf < L && buf[j + tf] == t) tf++;
                    uint64_t pos = cc[t]++;
                    out->post_doc[pos] = d;
                    out->post_tf[pos] = tf;
                    j += tf;
                }
            }
        }
    }
    return BM25X_OK;
}

extern "C" int bm25x_synth_queries(uint64_t seed, uint32_t nq, uint32_t vocab, uint32_t n_min, uint32_t n_max,
                                   double zipf_s, const uint64_t *post_off, uint32_t *q_off, uint32_t *q_terms) {
    if (!q_off || !q_terms || !post_off || n_min == 0 || n_max < n_min || vocab == 0) {
        bm25x_set_error("bm25x_synth_queries: bad arguments");
        return BM25X_ERR_INVALID;
    }
    std::vector<uint64_t> thrv;
    const uint64_t *thr = nullptr;
    if (zipf_s > 0.0) {
        zipf_thresholds(vocab, zipf_s, thrv);
        thr = thrv.data();
    }
    uint32_t pos = 0;
    q_off[0] = 0;
    std::vector<uint32_t> got;
    for (uint32_t i = 0; i < nq; i++) {
        uint32_t m = n_min;
        if (n_max > n_min) {
            uint64_t u = draw(seed, i, 0xFFFFFFFFu);
            m = n_min + (uint32_t)(((u >> 32) * (uint64_t)(n_max - n_min + 1)) >> 32);
        }
        got.clear();
        for (uint32_t j = 0; got.size() < m && j < 64 * m + 64; j++) {
            uint32_t t = draw_term(draw(seed, i, j), vocab, thr);
            if (post_off[t + 1] == post_off[t]) continue;
            if (std::find(got.begin(), got.end(), t) != got.end()) continue;
            got.push_back(t);
        }
        std::sort(got.begin(), got.end());
        for (uint32_t t : got) q_terms[pos++] = t;
        q_off[i + 1] = pos;
    }
    return BM25X_OK;
}
