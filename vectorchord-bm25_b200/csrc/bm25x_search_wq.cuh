// bm25x_search_wq.cuh — kernel v5: one WARP per query (sm_100a), for k <= 128 and <= 8 live terms.
//
// Every warp of the persistent grid is a complete, independent query engine: it fetches a query from the global work
// counter, plans its own doc-id chunks (lane j = term j; every term loads exactly quota*128 postings from where the
// previous window ended), streams them into its private single-buffered shared-memory stage with TMA bulk copies
// (cp.async.bulk + mbarrier; the other warps of the SM hide the load latency), unites the chunk's runs with its
// private tag map(s) (mark / test / resolve), re-scores the survivors of the f32 filter exactly in f64 and keeps its
// own candidate pool.  No CTA barrier, no producer or
// splitter warp, no shared pool: nothing a warp does depends on another warp.
//
// Exactness is the same argument as the CTA kernel: the f32 filter rejects only F < Sk·(1-2^-18)
// and exact-score ties by signature; everything else is ranked by (f64 score desc, doc id asc).
#pragma once

#include "bm25x_search_kernel.cuh"

namespace {

// Tuning knobs (compile time).  The defaults are the measured best of tools/time_variants.py on the 10M-doc corpus
// (profiles/README.md, "variants"): -1 = choose per class.
#ifndef BM25X_AU
#define BM25X_AU 4
#endif
#ifndef BM25X_BU
#define BM25X_BU 4
#endif
#ifndef BM25X_NSUB_SMALL
#define BM25X_NSUB_SMALL 1
#endif
#ifndef BM25X_LOG_S
#define BM25X_LOG_S -1
#endif
#ifndef BM25X_CBMUL
#define BM25X_CBMUL 2
#endif
#ifndef BM25X_SUBT
#define BM25X_SUBT 640
#endif
#ifndef BM25X_TWOMAP
#define BM25X_TWOMAP -1
#endif
#ifndef BM25X_EXACTQ
#define BM25X_EXACTQ 1
#endif
#ifndef BM25X_LAZYCUT
#define BM25X_LAZYCUT 1
#endif
#ifndef BM25X_BALANCED
#define BM25X_BALANCED 0
#endif
#ifndef BM25X_MAXWARPS
#define BM25X_MAXWARPS 16
#endif

template <int M_, int KP_>
struct WCfg {
    static constexpr int M = M_;                    // max live terms (lanes 0..M-1 own the terms)
    static constexpr int KP = KP_;                  // pool capacity (power of two >= k + LCAP)
    // block budget per chunk (Σ quota = CB exactly): two 128-posting blocks per term for m = M <= 4
    static constexpr int CB = BM25X_CBMUL * M_;
    // tag map bytes: queries of 4+ terms unite 1000+ postings per window and see far more slot collisions
    static constexpr int LOG_S = BM25X_LOG_S > 0 ? BM25X_LOG_S : (M_ >= 4 ? 13 : 12);
    static constexpr int SUB_TARGET = LOG_S >= 12 ? (BM25X_SUBT << (LOG_S - 12)) : (BM25X_SUBT >> (12 - LOG_S));  // postings per tag-map sub-window (classes with more than 4 terms)
    static constexpr int NSTG = 1;                  // stages per warp: 1 = rely on the other warps to hide the load latency
    static constexpr int LCAP = 64;                 // candidate / possible-duplicate list entries
    static constexpr int AU = BM25X_AU;             // mark phase: postings per lane and trip (independent loads in flight)
    static constexpr int BU = BM25X_BU;             // test phase: postings per lane and trip
    static constexpr int NSUB_SMALL = BM25X_NSUB_SMALL;  // sub-windows per chunk for classes up to 4 terms
    // Two half-size tag maps under independent hashes instead of one: a posting is a possible duplicate only when its
    // tag lost BOTH slots (a later run holding the same document overwrites both), which squares the false-alarm rate
    // (pays for its extra hash + byte load per posting only where collisions are frequent: the 4+ term classes;
    // measured: 3-term queries 22.8 ms single / 24.5 ms double, 1..8-term mix 66 ms single / 56 ms double)
    static constexpr bool TWOMAP = BM25X_TWOMAP >= 0 ? BM25X_TWOMAP != 0 : (M_ >= 4);
    static constexpr int LOG_M = TWOMAP ? LOG_S - 1 : LOG_S;
    // Loads of exactly quota*128 postings from where the last window ended (window end = doc id of the first posting
    // not loaded, one 4-byte read) instead of loads ending on block boundaries: no half-used blocks in the stage
    static constexpr bool EXACTQ = BM25X_EXACTQ != 0;
    static constexpr int STAGE_POSTINGS = EXACTQ ? CB * (int)BM25X_BLOCK  // every run: at most quota*128 postings
                                                 : (CB + M_) * (int)BM25X_BLOCK;  // + one partially consumed block per run
    static constexpr size_t stage_bytes = (size_t)STAGE_POSTINGS * sizeof(Posting);
    // per-warp shared memory
    static constexpr size_t off_stage = 0;
    static constexpr size_t off_map = off_stage + stage_bytes * NSTG;
    static constexpr size_t off_pool_s = off_map + ((size_t)1 << LOG_S);
    static constexpr size_t off_pool_d = off_pool_s + (size_t)KP * 8;
    static constexpr size_t off_pool_g = off_pool_d + (size_t)KP * 4;
    static constexpr size_t off_cand = off_pool_g + (size_t)KP * 4;
    static constexpr size_t off_dup = off_cand + (size_t)LCAP * 4;
    static constexpr size_t off_bar = off_dup + (size_t)LCAP * 4;
    static constexpr size_t warp_bytes = (off_bar + 8 * NSTG + 127) & ~(size_t)127;
    static constexpr size_t off_s1f = 0;  // CTA-shared: 1 KiB table first, then the warps
    static constexpr size_t shared_bytes = 1024;
    static constexpr int WARPS = (int)((227 * 1024 - shared_bytes) / warp_bytes) > BM25X_MAXWARPS
                                     ? BM25X_MAXWARPS
                                     : (int)((227 * 1024 - shared_bytes) / warp_bytes);
    static constexpr size_t total = shared_bytes + warp_bytes * WARPS;
    static constexpr int THREADS = WARPS * 32;
};

// One planned chunk: what every lane needs to issue its TMA copy and to process the chunk afterwards.
struct ChunkPlan {
    uint32_t lo, hi;     // doc window (hi already clamped to n_docs)
    uint32_t off, len;   // my run's placement inside the stage (postings); lanes >= m: len 0
    uint32_t gsrc;       // posting index (inside my term's list) of stage position `off`
    bool last;
};

template <class C>
struct WarpState {
    // query terms (lane j < m)
    uint32_t m, dfj, nb, quota_full;
    uint64_t pbase, bbase;
    float s0f;
    double s0d, ubd;
    // MaxScore pruning (warp-uniform): terms in ne_mask are no longer streamed; ub_ne = Σ of their score bounds
    uint32_t ne_mask;
    double ub_ne;
    // chunk planner: gpos = first posting of my term not yet consumed (exact, found by searching the landed chunk);
    // next_doc = its doc id when it has already been seen in shared memory (0 = unknown)
    uint32_t gpos, next_doc, lo, chunk;
};

// Tag-map slots of a document (see WCfg::TWOMAP) and the three questions asked of the map.
template <class C>
__device__ __forceinline__ uint32_t slot1(uint32_t doc) { return (doc * 0x9E3779B1u) >> (32 - C::LOG_M); }
template <class C>
__device__ __forceinline__ uint32_t slot2(uint32_t doc) { return (1u << C::LOG_M) + ((doc * 0x85EBCA77u) >> (32 - C::LOG_M)); }
template <class C>
__device__ __forceinline__ void tag_mark(uint8_t *map, uint32_t doc, uint8_t tag) {
    map[slot1<C>(doc)] = tag;
    if (C::TWOMAP) map[slot2<C>(doc)] = tag;
}
// true: no later run holds this document (the tag survived in a slot that every holder of the document writes)
template <class C>
__device__ __forceinline__ bool tag_intact(const uint8_t *map, uint32_t doc, uint32_t tag) {
    const uint32_t t1 = map[slot1<C>(doc)];
    if (!C::TWOMAP) return t1 == tag;
    const uint32_t t2 = map[slot2<C>(doc)];
    return (t1 == tag) | (t2 == tag);
}
// The last run holding the document either kept a slot — then it is the smaller of the two slot owners, later runs can
// only have taken the other one — or lost both and is listed as a possible duplicate itself.
template <class C>
__device__ __forceinline__ uint32_t tag_winner(const uint8_t *map, uint32_t doc) {
    const uint32_t t1 = map[slot1<C>(doc)];
    if (!C::TWOMAP) return t1 - 1u;
    const uint32_t t2 = map[slot2<C>(doc)];
    return min(t1, t2) - 1u;
}

// Plans the next chunk.  Loads start at the exact posting where the previous window ended (rounded down to the
// 16-byte TMA granule) and end on the block boundary chosen by the quota rule, so nothing is scanned twice.
// First half of the planner: issue the one global load the plan needs (first doc of the block just past my quota).
// Called as soon as gpos is known, so that the round trip hides behind the processing of the current chunk.
template <class C>
__device__ __forceinline__ uint32_t plan_prefetch(const SearchParams &p, const WarpState<C> &w, int lane) {
    const bool act = lane < (int)w.m && w.gpos < w.dfj && !((w.ne_mask >> lane) & 1u);
    const uint32_t quota = act ? (w.chunk < 2 ? 1u : w.quota_full) : 0u;
    uint32_t prop = INF;
    if (C::EXACTQ) {
        const uint32_t endp = (w.gpos & ~1u) + quota * BM25X_BLOCK;  // first posting past my load
        if (act && endp < w.dfj) prop = __ldg(&p.post[w.pbase + endp].doc);
    } else {
        const uint32_t ib = w.gpos / BM25X_BLOCK;
        if (act && ib + quota < w.nb) prop = __ldg(&p.blk[w.bbase + ib + quota].x);
    }
    return prop;
}

template <class C>
__device__ __forceinline__ ChunkPlan plan_chunk(const SearchParams &p, WarpState<C> &w, int lane, uint32_t prop) {
    const bool act = lane < (int)w.m && w.gpos < w.dfj && !((w.ne_mask >> lane) & 1u);
    const uint32_t quota = act ? (w.chunk < 2 ? 1u : w.quota_full) : 0u;
    const uint32_t ib = w.gpos / BM25X_BLOCK;
    // window end: the smallest "first doc of the block just past my quota" over the terms
    uint32_t hi = __reduce_min_sync(0xFFFFFFFFu, prop);
    if (w.chunk == 0 && hi != INF) hi = w.lo + max(1u, (hi - w.lo) >> 2);
    ChunkPlan c;
    c.len = 0;
    c.gsrc = w.gpos & ~1u;
    // a term whose next posting is known to lie at or past the window end has nothing in this chunk: do not load it
    // again (sparse terms next to dense ones would otherwise re-load the same block for thousands of chunks)
    if (act && !(hi != INF && w.next_doc >= hi)) {
        const uint32_t endp = min(C::EXACTQ ? c.gsrc + quota * BM25X_BLOCK : (ib + quota) * BM25X_BLOCK, w.dfj);
        c.len = (endp - c.gsrc + 1u) & ~1u;  // whole 16-byte units; an odd tail is the term's pad slot
    }
    uint32_t incl = c.len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    c.off = incl - c.len;
    c.lo = w.lo;
    c.hi = min(hi, p.n_docs);
    c.last = hi == INF;
    w.lo = hi;
    w.chunk++;
    return c;
}

template <class C>
__device__ __forceinline__ void issue_chunk(const SearchParams &p, const WarpState<C> &w, const ChunkPlan &c,
                                            uint8_t *stage, uint64_t *bar, int lane) {
    const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, c.len);
    // the stage was last read through the generic proxy by this warp: order those reads before the async-proxy writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(bar, total * (uint32_t)sizeof(Posting));
    __syncwarp();
    if (c.len > 0)
        tma_load_1d(stage + (size_t)c.off * sizeof(Posting), p.post + w.pbase + c.gsrc, c.len * (uint32_t)sizeof(Posting),
                    bar);
}

// Posting word of `doc` in a term that is not streamed any more (non-essential): block table first
// (SummaryTuple.{min,max}_document_id), then inside the 128-posting block.  ~24 dependent L2 reads; candidates only.
__device__ __forceinline__ uint32_t probe_global(const SearchParams &p, uint64_t pbase, uint64_t bbase, uint32_t nb,
                                                 uint32_t dfj, uint32_t doc) {
    uint32_t l = 0, r = nb;  // first block whose first doc is > doc
    while (l < r) {
        uint32_t mid = (l + r) >> 1;
        if (__ldg(&p.blk[bbase + mid].x) <= doc) l = mid + 1;
        else r = mid;
    }
    if (l == 0) return 0u;
    const uint32_t s = (l - 1) * BM25X_BLOCK, e = min(s + BM25X_BLOCK, dfj);
    const Posting *pp = p.post + pbase;
    l = s;
    r = e;
    while (l < r) {
        uint32_t mid = (l + r) >> 1;
        if (__ldg(&pp[mid].doc) < doc) l = mid + 1;
        else r = mid;
    }
    if (l < e && __ldg(&pp[l].doc) == doc) return __ldg(&pp[l].w);
    return 0u;
}

// Warp-private pool: (score bits, doc, signature), unsorted until pool_cut.
template <class C>
struct WPool {
    uint64_t *s;
    uint32_t *d, *g;
};

// Bitonic sort of the warp's pool (n2 = power of two >= n), best first; then keep the best `k`.
template <class C>
__device__ __forceinline__ void wpool_sort(const WPool<C> &pl, int n, int lane) {
    int n2 = 2;
    while (n2 < n) n2 <<= 1;
    for (int i = n + lane; i < n2; i += 32) {
        pl.s[i] = 0;
        pl.d[i] = INF;
        pl.g[i] = SIG_NONE;
    }
    __syncwarp();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = lane; i < (n2 >> 1); i += 32) {
                int a = 2 * i - (i & (stride - 1));
                int b = a + stride;
                uint64_t ka = pl.s[a], kb = pl.s[b];
                uint32_t da = pl.d[a], db = pl.d[b];
                bool desc = (a & size) == 0;
                bool sw = desc ? key_before(kb, db, ka, da) : key_before(ka, da, kb, db);
                if (sw) {
                    pl.s[a] = kb;
                    pl.s[b] = ka;
                    pl.d[a] = db;
                    pl.d[b] = da;
                    uint32_t ga = pl.g[a];
                    pl.g[a] = pl.g[b];
                    pl.g[b] = ga;
                }
            }
            __syncwarp();
        }
    }
}

struct WFilter {
    bool tv;
    float Flo;   // f32 scores below this cannot reach the top-k
    float ctf;   // lane j: single-term postings of run j pass iff tf >= ctf * s1[fn]   (the same test, solved for tf)
    double Sk;
    uint32_t dk, tie_sig, tie_dk;
};
__device__ __forceinline__ bool wfilter_pass(const WFilter &f, float F, uint32_t sig, uint32_t doc) {
    return F >= f.Flo && !(sig == f.tie_sig && doc > f.tie_dk);
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, 1) k_search_wq(const __grid_constant__ SearchParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    float *s1f = (float *)(smem + C::off_s1f);
    for (int i = threadIdx.x; i < 256; i += C::THREADS) s1f[i] = p.s1f[i];
    uint8_t *ws = smem + C::shared_bytes + C::warp_bytes * wid;
    uint8_t *map = ws + C::off_map;
    WPool<C> pl;
    pl.s = (uint64_t *)(ws + C::off_pool_s);
    pl.d = (uint32_t *)(ws + C::off_pool_d);
    pl.g = (uint32_t *)(ws + C::off_pool_g);
    uint32_t *cand = (uint32_t *)(ws + C::off_cand);
    uint32_t *dupl = (uint32_t *)(ws + C::off_dup);
    uint64_t *bars = (uint64_t *)(ws + C::off_bar);
    if (lane == 0) {
        for (int s = 0; s < C::NSTG; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    for (int i = lane; i < (1 << C::LOG_S) / 16; i += 32) ((uint4 *)map)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const uint32_t k = p.k;
    const double kEps = 1.0 / 262144.0;
    uint32_t parbits = 0;  // mbarrier phase parity per stage (bit s)

    for (;;) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(p.work_counter, 1);
        qi = __shfl_sync(0xFFFFFFFFu, qi, 0);
        if (qi >= (int)p.nq) break;
        const uint32_t qid = p.q_ids[qi];
        const uint32_t t0 = p.q_off[qi];
        WarpState<C> w;
        w.m = p.q_off[qi + 1] - t0;
        w.dfj = 0;
        w.nb = 0;
        w.pbase = w.bbase = 0;
        w.s0f = 0.f;
        w.s0d = 0.0;
        w.ubd = 0.0;
        w.ne_mask = 0u;
        w.ub_ne = 0.0;
        unsigned long long fetched = 0;
        if (lane < (int)w.m) {
            uint32_t term = p.q_terms[t0 + lane];
            w.dfj = p.df[term];
            w.pbase = p.post_off[term];
            w.bbase = p.blk_off[term];
            w.nb = (w.dfj + BM25X_BLOCK - 1) / BM25X_BLOCK;
            w.s0f = p.s0f[term];
            w.s0d = p.s0d[term];
            w.ubd = p.ubd[term];
        }
        {
            uint64_t sumdf = w.dfj;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sumdf += __shfl_xor_sync(0xFFFFFFFFu, sumdf, o);
            // one block each, the rest of the budget ∝ df, the rounding leftover to the first terms
            const uint32_t extra = (uint32_t)C::CB - w.m;
            const uint32_t share = lane < (int)w.m ? (uint32_t)(((uint64_t)extra * w.dfj) / sumdf) : 0u;
            const uint32_t left = extra - __reduce_add_sync(0xFFFFFFFFu, share);
#if BM25X_BALANCED
            // largest-remainder split: the leftover blocks go to the terms that lost most by the flooring, so equally
            // frequent terms get equal loads (not yet the default: unmeasured, see DESIGN.md §4.1 "known inefficiency")
            const uint64_t rem = lane < (int)w.m ? ((uint64_t)extra * w.dfj) % sumdf : 0ull;
            uint32_t rank = 0;
            for (uint32_t i = 0; i < w.m; ++i) {
                const uint64_t ri = __shfl_sync(0xFFFFFFFFu, rem, i);
                rank += (ri > rem || (ri == rem && (int)i < lane)) ? 1u : 0u;
            }
            w.quota_full = lane < (int)w.m ? 1u + share + (rank < left ? 1u : 0u) : 0u;
#else
            w.quota_full = lane < (int)w.m ? 1u + share + (lane < (int)left ? 1u : 0u) : 0u;
#endif
        }
        w.gpos = 0;
        w.next_doc = 0;
        w.lo = 0;
        w.chunk = 0;
        const uint32_t m = w.m;
        // per-query pool / threshold state (warp-uniform registers)
        int pn = 0;
        WFilter f;
        f.tv = false;
        f.Flo = -1.f;
        f.Sk = 0.0;
        f.dk = INF;
        f.tie_sig = SIG_NONE;
        f.tie_dk = INF;
        f.ctf = 0.f;  // no threshold yet: everything passes
        bool dense_query = false;  // sticky: a chunk of this query needed the dense-overlap path

        // cut the pool back to k and refresh the threshold
        auto pool_cut = [&]() {
            wpool_sort<C>(pl, pn, lane);
            // drop duplicates (the dense-overlap pass may emit a document twice: equal (score, doc) are adjacent)
            {
                int out = 0;
                for (int base = 0; base < pn; base += 32) {
                    const int i = base + lane;
                    uint64_t sv = 0;
                    uint32_t dv = INF, gv = SIG_NONE;
                    bool keep = false;
                    if (i < pn) {
                        sv = pl.s[i];
                        dv = pl.d[i];
                        gv = pl.g[i];
                        keep = i == 0 || !(pl.s[i - 1] == sv && pl.d[i - 1] == dv);
                    }
                    const uint32_t mk = __ballot_sync(0xFFFFFFFFu, keep);
                    __syncwarp();
                    if (keep) {
                        const int o = out + __popc(mk & lt_mask);
                        pl.s[o] = sv;
                        pl.d[o] = dv;
                        pl.g[o] = gv;
                    }
                    out += __popc(mk);
                    __syncwarp();
                }
                pn = out;
            }
            if (pn > (int)k) pn = (int)k;
            if (pn == (int)k) {
                f.Sk = __longlong_as_double((long long)pl.s[k - 1]);
                f.dk = pl.d[k - 1];
                f.tie_sig = pl.g[k - 1];
                f.tie_dk = f.tie_sig != SIG_NONE ? f.dk : INF;
                // MaxScore: move the terms with the smallest score bounds out of the streamed set while the sum of
                // their bounds stays below 5 % of the k-th score (a document holding only such terms cannot enter; for
                // the others the bound is added back in the filter and the exact contribution is probed in phase D)
                if (p.prune) {
                    for (;;) {
                        const bool ess = lane < (int)m && !((w.ne_mask >> lane) & 1u);
                        unsigned long long key = ess ? (unsigned long long)__double_as_longlong(w.ubd) : ~0ull;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
                            key = other < key ? other : key;
                        }
                        const uint32_t who = __ballot_sync(0xFFFFFFFFu, ess && (unsigned long long)__double_as_longlong(w.ubd) == key);
                        const uint32_t ness = __popc(__ballot_sync(0xFFFFFFFFu, ess));
                        if (who == 0u || ness <= 1u) break;
                        const double ub = __longlong_as_double((long long)key);
                        if (!(w.ub_ne + ub <= 0.05 * f.Sk)) break;
                        w.ne_mask |= 1u << (__ffs(who) - 1);
                        w.ub_ne += ub;
                    }
                    if (w.ne_mask) f.tie_dk = INF;  // a single-term document may hold non-streamed terms: no tie shortcut
                }
                const double flo = f.Sk * (1.0 - kEps) - w.ub_ne;
                f.Flo = __double2float_rd(flo);
                // F = s0·tf/(tf+s1) >= flo  ⇔  tf >= flo/(s0-flo)·s1  (s0 > flo), never when s0 <= flo.  Solved in f64
                // from the exact s0 (no cancellation trouble), shrunk by 2^-20 to stay conservative in f32.
                f.ctf = __int_as_float(0x7f800000);  // +inf
                if (lane < (int)m && w.s0d > flo) f.ctf = __double2float_rd(flo / (w.s0d - flo) * (1.0 - 1.0 / 1048576.0));
                f.tv = true;
            }
        };

        int stage = 0;
        // ---- prime the pipeline: plan + issue chunk 0 ----
        ChunkPlan cur = plan_chunk<C>(p, w, lane, plan_prefetch<C>(p, w, lane));
        issue_chunk<C>(p, w, cur, ws + C::off_stage + C::stage_bytes * stage, &bars[stage], lane);
        for (;;) {
            // ---- wait for the current chunk ----
            mbar_wait(&bars[stage], (parbits >> stage) & 1u);
            parbits ^= 1u << stage;
            const Posting *st = (const Posting *)(ws + C::off_stage + C::stage_bytes * stage);
            const uint32_t clo = cur.lo, chi = cur.hi;
            fetched += cur.len;
            // exact in-window range of my run (lane < m): [run_a, run_e) — one binary search for the window end; the
            // load started at most one posting before the window start
            uint32_t run_a = cur.off, run_e = cur.off;
            if (cur.len > 0) {
                if (st[run_a].doc < clo) run_a++;
                uint32_t l = run_a, r = cur.off + cur.len;
                while (l < r) {
                    uint32_t mid = (l + r) >> 1;
                    if (st[mid].doc < chi) l = mid + 1;
                    else r = mid;
                }
                run_e = l;
                w.gpos = cur.gsrc + (run_e - cur.off);  // first posting of my term at or past the window end
                w.next_doc = run_e < cur.off + cur.len ? st[run_e].doc : 0u;
            }
            // ---- prefetch: plan + issue the next chunk into the other stage (overlaps the processing below) ----
            ChunkPlan nxt;
            nxt.last = true;
            nxt.len = nxt.off = 0;
            nxt.lo = nxt.hi = nxt.gsrc = 0;
            const bool have_next = !cur.last;
            uint32_t prop_next = INF;
            if (have_next) prop_next = plan_prefetch<C>(p, w, lane);  // consumed after the processing below (NSTG == 1)
            if (C::NSTG == 2 && have_next) {
                nxt = plan_chunk<C>(p, w, lane, prop_next);
                issue_chunk<C>(p, w, nxt, ws + C::off_stage + C::stage_bytes * (stage ^ 1), &bars[stage ^ 1], lane);
            }

            // The chunk is processed in doc sub-windows of at most ~SUB_TARGET postings: loads can be large while the
            // tag map only ever has to unite a few hundred postings.
            // (classes up to 4 terms load two blocks per term: one sub-window, resolved at compile time)
            uint32_t nsub = (uint32_t)C::NSUB_SMALL;
            if (C::M > 4) {
                const uint32_t chunk_postings = __reduce_add_sync(0xFFFFFFFFu, run_e - run_a);
                nsub = max(1u, (chunk_postings + C::SUB_TARGET - 1) / C::SUB_TARGET);
            }
            uint32_t sub_next = run_a;
            for (uint32_t sub = 0; sub < nsub; ++sub) {
            const uint32_t lo = clo + (uint32_t)(((uint64_t)(chi - clo) * sub) / nsub);
            const uint32_t hi = sub + 1 == nsub ? chi : clo + (uint32_t)(((uint64_t)(chi - clo) * (sub + 1)) / nsub);
            const uint32_t my_a = sub_next;
            uint32_t my_e = run_e;
            if (sub + 1 < nsub) {  // my run's end inside this sub-window
                uint32_t l = my_a, r = run_e;
                while (l < r) {
                    uint32_t mid = (l + r) >> 1;
                    if (st[mid].doc < hi) l = mid + 1;
                    else r = mid;
                }
                my_e = l;
            }
            sub_next = my_e;
            uint32_t nd = 0, nc = 0;  // list lengths (warp-uniform)
            // exact re-score of the listed candidates → pool
            auto flush = [&]() {
                for (uint32_t base = 0; base < nc; base += 32) {
                    const bool has = base + lane < nc;
                    const uint32_t ent = has ? cand[base + lane] : 0u;
                    // entry flavours: (run << 16 | stage position) from the tag-map phases; bit 31 set: no twin rule;
                    // bits 31+30 set: document given as offset from the window start (dense accumulator path)
                    const uint32_t doc = (ent >> 30) == 3u ? lo + (ent & 0x3FFFFFFFu) : st[ent & 0xFFFFu].doc;
                    double Sx = 0.0;
                    uint32_t cnt = 0, cnt_streamed = 0, sig = SIG_NONE;
                    for (uint32_t jj = 0; jj < m; ++jj) {
                        const uint32_t a = __shfl_sync(0xFFFFFFFFu, my_a, jj), e = __shfl_sync(0xFFFFFFFFu, my_e, jj);
                        const double s0d = __shfl_sync(0xFFFFFFFFu, w.s0d, jj);
                        const bool ne = (w.ne_mask >> jj) & 1u;  // not streamed: look the posting up in HBM
                        uint64_t pb = 0, bb = 0;
                        uint32_t nbj = 0, dfq = 0;
                        if (ne) {
                            pb = __shfl_sync(0xFFFFFFFFu, w.pbase, jj);
                            bb = __shfl_sync(0xFFFFFFFFu, w.bbase, jj);
                            nbj = __shfl_sync(0xFFFFFFFFu, w.nb, jj);
                            dfq = __shfl_sync(0xFFFFFFFFu, w.dfj, jj);
                        }
                        if (!has) continue;
                        const uint32_t wv = ne ? probe_global(p, pb, bb, nbj, dfq, doc) : find_in(st, a, e, doc);
                        if (wv) {
                            Sx = __dadd_rn(Sx, score_f64(wv, s0d, p.s1d));
                            cnt++;
                            cnt_streamed += ne ? 0u : 1u;
                            sig = make_sig(jj, wv);
                        }
                    }
                    bool keep = has;
                    if (keep && p.allow && !((p.allow[doc >> 3] >> (doc & 7u)) & 1u)) keep = false;
                    // a posting whose tag won its slot although other runs hold the document: its twin carries it
                    // (only runs that are streamed take part in the tag map: probed terms do not make a twin)
                    if (keep && cnt_streamed > 1 && (ent >> 31) == 0 && tag_intact<C>(map, doc, ((ent >> 16) & 0x7FFFu) + 1u))
                        keep = false;
                    keep = keep && (!f.tv || Sx > f.Sk || (Sx == f.Sk && doc < f.dk));
                    const uint32_t mk = __ballot_sync(0xFFFFFFFFu, keep);
                    if (keep) {
                        const int idx = pn + __popc(mk & lt_mask);
                        pl.s[idx] = (uint64_t)__double_as_longlong(Sx);
                        pl.d[idx] = doc;
                        pl.g[idx] = cnt == 1 ? sig : SIG_NONE;
                    }
                    pn += __popc(mk);
                    __syncwarp();
                    // Re-sorting the pool is the expensive part of a large k (bitonic sort of KP entries): once a
                    // threshold exists, the big pool is cut only when it is about to overflow.
                    const bool lazy = BM25X_LAZYCUT && C::KP > 128 && f.tv;
                    if (pn > C::KP - 32 || (!lazy && pn >= (int)k + 32)) pool_cut();
                }
                nc = 0;
            };
            auto push_cand = [&](bool c, uint32_t ent) {
                const uint32_t mc = __ballot_sync(0xFFFFFFFFu, c);
                if (mc) {
                    if (c) cand[nc + __popc(mc & lt_mask)] = ent;
                    nc += __popc(mc);
                    __syncwarp();
                    if (nc > (uint32_t)C::LCAP - 32) flush();
                }
            };

            // Dense-overlap chunks (head terms: most documents hold several query terms) have a narrow doc window:
            // when it fits, scores are summed in a dense f32 accumulator indexed by doc - lo (in the tag-map memory).
            constexpr uint32_t ACC_DOCS = (1u << C::LOG_S) / 4u;
            const bool can_acc = hi - lo <= ACC_DOCS;
            bool dense = dense_query && can_acc;  // the previous chunk was dense: skip the tag-map phases right away
            // ---- A: mark (a single-term query has nothing to unite: no tag map at all) ----
            const bool multi = m > 1;
            if (!dense && multi) {
                for (uint32_t j = 0; j < m; ++j) {
                    const uint32_t a = __shfl_sync(0xFFFFFFFFu, my_a, j), e = __shfl_sync(0xFFFFFFFFu, my_e, j);
                    const uint8_t tagv = (uint8_t)(j + 1);
                    // AU postings per lane and trip while whole trips remain (independent loads in flight), then singly
                    uint32_t i = a + lane;
                    if (C::AU > 1) {
                        for (; i + 32u * (C::AU - 1) < e; i += 32 * C::AU) {
                            uint32_t dd[C::AU];
#pragma unroll
                            for (int u = 0; u < C::AU; ++u) dd[u] = st[i + 32u * u].doc;
#pragma unroll
                            for (int u = 0; u < C::AU; ++u) tag_mark<C>(map, dd[u], tagv);
                        }
                    }
                    for (; i < e; i += 32) tag_mark<C>(map, st[i].doc, tagv);
                }
                __syncwarp();
            }
            // ---- B: test; score + filter the singles; list the possible duplicates (BU postings per lane) ----
            for (uint32_t j = 0; j < m && !dense; ++j) {
                const uint32_t a = __shfl_sync(0xFFFFFFFFu, my_a, j), e = __shfl_sync(0xFFFFFFFFu, my_e, j);
                const float ctf = __shfl_sync(0xFFFFFFFFu, f.ctf, j);
                const uint32_t tagv = j + 1u;
                for (uint32_t base = a; base < e; base += 32 * C::BU) {
                    uint32_t idx[C::BU];
                    Posting pp[C::BU];
                    bool vv[C::BU], dd[C::BU], gg[C::BU];
#pragma unroll
                    for (int u = 0; u < C::BU; ++u) {
                        idx[u] = base + lane + 32u * u;
                        vv[u] = idx[u] < e;
                        pp[u] = st[vv[u] ? idx[u] : a];
                    }
                    bool anyd = false, anyg = false;
#pragma unroll
                    for (int u = 0; u < C::BU; ++u) {
                        dd[u] = vv[u] && multi && !tag_intact<C>(map, pp[u].doc, tagv);
                        // threshold test in the tf domain (no division), unconditional: no branches; the signature
                        // tie rule is only evaluated for the few postings that get this far
                        const float r = ctf * s1f[pp[u].w & 0xFFu];
                        gg[u] = ((float)(pp[u].w >> 8) >= r) & vv[u] & !dd[u];
                        anyd |= dd[u];
                        anyg |= gg[u];
                    }
                    if (__any_sync(0xFFFFFFFFu, anyd)) {
                        uint32_t q = nd;
#pragma unroll
                        for (int u = 0; u < C::BU; ++u) {
                            const uint32_t md = __ballot_sync(0xFFFFFFFFu, dd[u]);
                            const uint32_t qq = q + __popc(md & lt_mask);
                            if (dd[u] && qq < (uint32_t)C::LCAP) dupl[qq] = (j << 16) | idx[u];
                            q += __popc(md);
                        }
                        nd = q;
                        if (nd > (uint32_t)C::LCAP) {
                            dense = true;
                            break;
                        }
                    }
                    if (__any_sync(0xFFFFFFFFu, anyg)) {
#pragma unroll
                        for (int u = 0; u < C::BU; ++u) {
                            const bool c = gg[u] && !(make_sig(j, pp[u].w) == f.tie_sig && pp[u].doc > f.tie_dk);
                            push_cand(c, (j << 16) | idx[u]);
                        }
                    }
                }
            }
            __syncwarp();
            if (!dense) {
                // ---- C: resolve the possible duplicates; exactly one emitter per document ----
                if (nd <= 32) {
                    // Every run that holds the document but did not win its slot is in the list, so the list itself
                    // tells which listed postings belong together (MATCH.ANY); only the slot's winner run — whose
                    // posting, if it is the same document, stayed silent — has to be searched.
                    const bool has = lane < (int)nd;
                    const uint32_t ent = has ? dupl[lane] : 0u;
                    const uint32_t j = ent >> 16;
                    const Posting v = st[ent & 0xFFFFu];
                    const unsigned long long key = has ? (unsigned long long)v.doc : ((1ull << 32) | (unsigned)lane);
                    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
                    const uint32_t winner = has ? tag_winner<C>(map, v.doc) : 0u;
                    const float Fm = score_f32(v.w, __shfl_sync(0xFFFFFFFFu, w.s0f, j & 31u), s1f);
                    float F = 0.f;
                    uint32_t cnt = 0, rem = peers;
                    while (__any_sync(0xFFFFFFFFu, rem != 0u)) {  // entries are listed in ascending run order
                        const int src = rem ? __ffs(rem) - 1 : lane;
                        const float val = __shfl_sync(0xFFFFFFFFu, Fm, src);
                        if (rem) {
                            F += val;
                            cnt++;
                            rem &= rem - 1u;
                        }
                    }
                    const bool owner = has && lane == __ffs(peers) - 1;
                    const uint32_t wa = __shfl_sync(0xFFFFFFFFu, my_a, winner & 31u), we = __shfl_sync(0xFFFFFFFFu, my_e, winner & 31u);
                    const float ws0 = __shfl_sync(0xFFFFFFFFu, w.s0f, winner & 31u);
                    if (owner) {
                        const uint32_t wv = find_in(st, wa, we, v.doc);
                        if (wv) {
                            F += score_f32(wv, ws0, s1f);
                            cnt++;
                        }
                    }
                    push_cand(owner && wfilter_pass(f, F, cnt == 1 ? make_sig(j, v.w) : SIG_NONE, v.doc), ent);
                } else
                for (uint32_t base = 0; base < nd; base += 32) {
                    const bool has = base + lane < nd;
                    const uint32_t ent = has ? dupl[base + lane] : 0u;
                    const uint32_t j = ent >> 16;
                    const Posting v = st[ent & 0xFFFFu];
                    const uint32_t winner = tag_winner<C>(map, v.doc);
                    float F = 0.f;
                    uint32_t cnt = 0;
                    bool owner = has;
                    for (uint32_t jj = 0; jj < m; ++jj) {
                        const uint32_t a = __shfl_sync(0xFFFFFFFFu, my_a, jj), e = __shfl_sync(0xFFFFFFFFu, my_e, jj);
                        const float s0 = __shfl_sync(0xFFFFFFFFu, w.s0f, jj);
                        if (!owner) continue;
                        const uint32_t wv = jj == j ? v.w : find_in(st, a, e, v.doc);
                        if (!wv) continue;
                        if (jj < j && jj != winner) {  // a lower run also detected this document: it emits
                            owner = false;
                            continue;
                        }
                        F += score_f32(wv, s0, s1f);
                        cnt++;
                    }
                    push_cand(owner && wfilter_pass(f, F, cnt == 1 ? make_sig(j, v.w) : SIG_NONE, v.doc), ent);
                }
            } else if (can_acc) {
                // ---- dense overlap, narrow window: dense accumulator ----
                // Candidates listed so far carry tag-map semantics: settle them first.  Documents already emitted by B
                // for this chunk are emitted again here; pool_cut() removes the duplicates (same doc ⇒ same exact score
                // ⇒ adjacent after the sort).
                flush();
                dense_query = true;
                float *acc = (float *)map;
                const uint32_t span = hi - lo;
                for (uint32_t i = lane; i < span; i += 32) acc[i] = 0.f;
                __syncwarp();
                for (uint32_t j = 0; j < m; ++j) {  // docs are distinct inside a run: no write conflicts
                    const uint32_t a = __shfl_sync(0xFFFFFFFFu, my_a, j), e = __shfl_sync(0xFFFFFFFFu, my_e, j);
                    const float s0 = __shfl_sync(0xFFFFFFFFu, w.s0f, j);
                    for (uint32_t i = a + lane; i < e; i += 32) {
                        const Posting v = st[i];
                        acc[v.doc - lo] += score_f32(v.w, s0, s1f);
                    }
                    __syncwarp();
                }
                for (uint32_t base = 0; base < span; base += 32) {
                    const uint32_t o = base + lane;
                    const float F = o < span ? acc[o] : 0.f;
                    if (__any_sync(0xFFFFFFFFu, F > 0.f && F >= f.Flo)) push_cand(F > 0.f && F >= f.Flo, 0xC0000000u | o);
                }
            } else {
                // ---- dense overlap, wide window (rare): every in-window posting looks its document up in the other
                // runs; the posting of the lowest run holding the document emits it (no tag-map twin rule) ----
                flush();
                for (uint32_t j = 0; j < m; ++j) {
                    const uint32_t a = __shfl_sync(0xFFFFFFFFu, my_a, j), e = __shfl_sync(0xFFFFFFFFu, my_e, j);
                    for (uint32_t base = a; base < e; base += 32) {
                        const uint32_t i = base + lane;
                        Posting v = st[i < e ? i : a];
                        const bool valid = i < e;
                        float F = 0.f;
                        uint32_t cnt = 0;
                        bool owner = valid;
                        for (uint32_t jj = 0; jj < m; ++jj) {
                            const uint32_t aa = __shfl_sync(0xFFFFFFFFu, my_a, jj), ee = __shfl_sync(0xFFFFFFFFu, my_e, jj);
                            const float s0 = __shfl_sync(0xFFFFFFFFu, w.s0f, jj);
                            if (!owner) continue;
                            const uint32_t wv = jj == j ? v.w : find_in(st, aa, ee, v.doc);
                            if (!wv) continue;
                            if (jj < j) {
                                owner = false;
                                continue;
                            }
                            F += score_f32(wv, s0, s1f);
                            cnt++;
                        }
                        push_cand(owner && wfilter_pass(f, F, cnt == 1 ? make_sig(j, v.w) : SIG_NONE, v.doc),
                                  0x80000000u | (j << 16) | i);
                    }
                }
            }
            if (nc) flush();
            // zero the tag map for the next chunk
            __syncwarp();
            if (multi)
                for (int i = lane; i < (1 << C::LOG_S) / 16; i += 32) ((uint4 *)map)[i] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            }  // sub-windows
            if (!have_next) break;
            if (C::NSTG == 1) {  // single-buffered: the stage is free again only now
                nxt = plan_chunk<C>(p, w, lane, prop_next);
                issue_chunk<C>(p, w, nxt, ws + C::off_stage + C::stage_bytes * stage, &bars[stage], lane);
            } else {
                stage ^= 1;
            }
            cur = nxt;
        }
        // ---- Results::into_sorted_vec (search.rs:281) ----
        if (pn > 0) pool_cut();
        const size_t obase = (size_t)qid * k;
        for (uint32_t i = lane; i < k; i += 32) {
            uint32_t d = INF;
            double sc = 0.0;
            if ((int)i < pn) {
                d = pl.d[i];
                sc = __longlong_as_double((long long)pl.s[i]);
            }
            p.out_doc[obase + i] = d;
            p.out_score[obase + i] = (float)sc;
            if (p.out_score64) p.out_score64[obase + i] = sc;
            if (p.out_payload) {
                uint16_t a = 0, b = 0, cc = 0;
                if ((int)i < pn) {
                    a = p.payload[(size_t)d * 3 + 0];
                    b = p.payload[(size_t)d * 3 + 1];
                    cc = p.payload[(size_t)d * 3 + 2];
                }
                p.out_payload[(obase + i) * 3 + 0] = a;
                p.out_payload[(obase + i) * 3 + 1] = b;
                p.out_payload[(obase + i) * 3 + 2] = cc;
            }
        }
        if (lane == 0) p.out_n[qid] = (uint32_t)pn;
        if (p.fetched) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) fetched += __shfl_xor_sync(0xFFFFFFFFu, fetched, o);
            if (lane == 0) atomicAdd(p.fetched, fetched);
        }
        __syncwarp();
    }
}

}  // namespace
