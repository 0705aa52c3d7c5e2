/* bm25_synth.c — bulk form of the oracle's synthetic-corpus generator (TEST INFRASTRUCTURE, see bm25_oracle.h).
 *
 * Same spec as orc_synth_doc (mirrors the reference's fuzz generator, tests/fuzz:168-205: draw L term ids
 * with replacement, aggregate duplicates into (term, tf), len = L; doc ids = generation order), produced for a whole
 * corpus at once as the term-major CSR the reference's flush() sees (Mapping(key, doc, tf) sorted by (key, doc),
 * crates/bm25/src/segment.rs:19-45).  bench.py --impl reference uses it so that the reference arm never touches the
 * product library; tests compare it bit for bit with the product's generator. */
#include <omp.h>
#include <stdlib.h>
#include <string.h>

#include "bm25_oracle.h"

/* Pass structure: the documents are cut into `nchunk` contiguous ranges; pass 1 counts postings per (chunk, term),
 * a prefix sum gives every chunk its private write cursor per term, pass 2 regenerates the documents and writes.
 * Doc ids ascend inside a chunk and chunks are in doc order, so every term's postings come out doc-ascending. */
int orc_synth_corpus(uint64_t seed, uint32_t n_docs, uint32_t vocab, uint32_t len_min, uint32_t len_max, double zipf_s,
                     int nthreads, uint32_t *doc_len /* [n_docs] */, uint64_t *post_off /* [vocab+1] */,
                     uint32_t *post_doc, uint32_t *post_tf, uint64_t cap /* capacity of post_doc / post_tf */,
                     uint64_t *n_post_out) {
    if (nthreads < 1) nthreads = 1;
    int nchunk = nthreads;
    if ((uint32_t)nchunk > n_docs) nchunk = n_docs ? (int)n_docs : 1;
    uint64_t *thr = NULL;
    if (zipf_s > 0.0) {
        thr = (uint64_t *)malloc(sizeof(uint64_t) * vocab);
        if (!thr) return -1;
        orc_zipf_thresholds(vocab, zipf_s, thr);
    }
    uint64_t *cnt = (uint64_t *)calloc((size_t)nchunk * vocab, sizeof(uint64_t));
    if (!cnt) {
        free(thr);
        return -1;
    }
    const uint32_t lmax = len_max > len_min ? len_max : len_min;
    int fail = 0;
#pragma omp parallel num_threads(nthreads)
    {
        uint32_t *terms = (uint32_t *)malloc(sizeof(uint32_t) * (lmax ? lmax : 1));
        uint32_t *tfs = (uint32_t *)malloc(sizeof(uint32_t) * (lmax ? lmax : 1));
        if (!terms || !tfs) {
#pragma omp atomic write
            fail = 1;
        }
#pragma omp barrier
        if (!fail) {
#pragma omp for schedule(static, 1)
            for (int c = 0; c < nchunk; c++) {
                const uint32_t d0 = (uint32_t)((uint64_t)n_docs * c / nchunk), d1 = (uint32_t)((uint64_t)n_docs * (c + 1) / nchunk);
                uint64_t *mine = cnt + (size_t)c * vocab;
                for (uint32_t d = d0; d < d1; d++) {
                    uint32_t L = 0;
                    const int n = orc_synth_doc(seed, d, vocab, len_min, len_max, thr, terms, tfs, &L);
                    doc_len[d] = L;
                    for (int i = 0; i < n; i++) mine[terms[i]]++;
                }
            }
            /* prefix: post_off, then per-chunk cursors (cnt becomes the start of each chunk's slice) */
#pragma omp single
            {
                uint64_t run = 0;
                for (uint32_t t = 0; t < vocab; t++) {
                    post_off[t] = run;
                    for (int c = 0; c < nchunk; c++) {
                        const uint64_t x = cnt[(size_t)c * vocab + t];
                        cnt[(size_t)c * vocab + t] = run;
                        run += x;
                    }
                }
                post_off[vocab] = run;
                *n_post_out = run;
                if (run > cap) fail = 2;
            }
            if (!fail) {
#pragma omp for schedule(static, 1)
                for (int c = 0; c < nchunk; c++) {
                    const uint32_t d0 = (uint32_t)((uint64_t)n_docs * c / nchunk), d1 = (uint32_t)((uint64_t)n_docs * (c + 1) / nchunk);
                    uint64_t *cur = cnt + (size_t)c * vocab;
                    for (uint32_t d = d0; d < d1; d++) {
                        uint32_t L = 0;
                        const int n = orc_synth_doc(seed, d, vocab, len_min, len_max, thr, terms, tfs, &L);
                        for (int i = 0; i < n; i++) {
                            const uint64_t at = cur[terms[i]]++;
                            post_doc[at] = d;
                            post_tf[at] = tfs[i];
                        }
                    }
                }
            }
        }
        free(terms);
        free(tfs);
    }
    free(cnt);
    free(thr);
    return fail ? -fail : 0;
}

/* Queries per SURVEY §8d: query i draws nterms(i) in [nmin, nmax] (draw 0xFFFFFFFF), then distinct terms with df > 0
 * from the corpus distribution (draws 0, 1, ...; at most 64*m + 64 attempts), sorted ascending.
 * q_off[nq+1]; q_terms must hold nq * nmax entries.  df(t) = post_off[t+1] - post_off[t]. */
int orc_synth_queries(uint64_t seed, uint32_t nq, uint32_t vocab, uint32_t nmin, uint32_t nmax, double zipf_s,
                      const uint64_t *post_off, uint32_t *q_off, uint32_t *q_terms) {
    uint64_t *thr = NULL;
    if (zipf_s > 0.0) {
        thr = (uint64_t *)malloc(sizeof(uint64_t) * vocab);
        if (!thr) return -1;
        orc_zipf_thresholds(vocab, zipf_s, thr);
    }
    uint32_t out = 0;
    q_off[0] = 0;
    for (uint32_t i = 0; i < nq; i++) {
        uint32_t m = nmin;
        if (nmax > nmin) {
            const uint64_t u = orc_draw(seed, i, 0xFFFFFFFFu);
            m = nmin + (uint32_t)(((u >> 32) * (uint64_t)(nmax - nmin + 1)) >> 32);
        }
        uint32_t *got = q_terms + out;
        uint32_t ng = 0;
        for (uint32_t j = 0; ng < m && j < 64 * m + 64; j++) {
            const uint32_t t = orc_draw_term(orc_draw(seed, i, j), vocab, thr);
            if (post_off[t + 1] == post_off[t]) continue;
            int dup = 0;
            for (uint32_t x = 0; x < ng; x++) dup |= got[x] == t;
            if (!dup) got[ng++] = t;
        }
        for (uint32_t a = 1; a < ng; a++) { /* insertion sort: m is tiny */
            const uint32_t v = got[a];
            uint32_t b = a;
            while (b > 0 && got[b - 1] > v) {
                got[b] = got[b - 1];
                b--;
            }
            got[b] = v;
        }
        out += ng;
        q_off[i + 1] = out;
    }
    free(thr);
    return 0;
}
