// bm25x_search_kernel.cuh — the sm_100a search kernel (device code only; host driver in bm25x_search.cu).
//
// One persistent CTA = 1 producer warp + 1 splitter warp + W merge warps.  Per query the producer cuts the doc-id
// space into chunks whose postings fit one shared-memory stage and streams them in with TMA bulk copies
// (cp.async.bulk + mbarrier, STAGES-deep ring).  The splitter cuts every landed chunk into W doc-id sub-windows
// (one binary search per run and boundary) so that each merge warp owns all postings of its documents.  A merge
// warp turns its sub-window into scored documents without talking to the other warps:
//
//   hot path (lane-parallel, warp-private tag map, no CTA barrier inside — kernel v3):
//     A  mark   every posting writes its run tag (j+1) into a 32 KiB byte map at hash(doc)         (plain st.shared.u8)
//     B  test   every posting re-reads its slot: tag still mine  → no other run touched the slot → the document is
//               (very probably) single-term: f32 score, threshold filter, done;
//               tag differs → "possible duplicate" (true multi-term document or a hash collision) → dup list
//     C  dups   each dup-list posting binary-searches the other runs; exactly one posting per document (the lowest
//               run that detected the clash) emits the document with its full f32 score
//     D  exact  the few survivors of the filter are re-scored in f64 in the reference's operation order
//               (Cache::evaluate, bm25.rs:355-358, summed over ascending terms) and enter the candidate pool;
//               a posting whose tag "won" its slot although other runs hold the same document is dropped here
//               (its detecting twin carries the document)
//   cold path (kernel v1): equal-width doc-id buckets + per-thread m-way merge; used when the dup list or the
//     candidate queue overflows (dense-overlap chunks, adversarial score order).
//
// Exactness: the f32 filter only rejects a document when its f32 score is below Sk·(1-2^-18), Sk = exact f64 k-th
// best so far; the f32 error bound is < 2^-18 relative.  Survivors are ranked by exact f64 score,
// ties by ascending doc id.
#pragma once

#include "bm25x_common.h"

// One launch = the queries of one term-count class (shared by the translation units of the library).
struct SearchParams {
    const Posting *post;
    const uint64_t *post_off;
    const uint32_t *df;
    const uint64_t *blk_off;
    const uint2 *blk;
    const float *blk_ub;                // [n_blocks] per-block score bound (SummaryTuple.wand_*)
    const float *s0f;
    const double *s0d;
    const double *s1d;
    const float *s1f;
    const uint16_t *payload;
    const double *ubd;                  // per-term upper bound of one posting's exact score
    unsigned long long *fetched;        // Σ postings actually loaded into shared memory (pruning statistics)
    uint8_t *pool_scratch;              // k > 1024: per-warp candidate pools in HBM (k_search_ring, RCfg::POOL_GLOBAL)
    int prune;
    float s1f_min;                      // min over the documents of s1f[fieldnorm]
    uint32_t n_docs;
    // one launch = the queries of one term-count class
    const uint32_t *q_ids;    // original query index
    const uint32_t *q_off;    // [nq+1]
    const uint32_t *q_terms;  // canonical: ascending, distinct, df > 0
    uint32_t nq;
    uint32_t k;
    const uint8_t *allow;
    int *work_counter;
    uint32_t *out_doc;
    float *out_score;
    double *out_score64;
    uint16_t *out_payload;
    uint32_t *out_n;
};

namespace {

constexpr uint32_t INF = BM25X_DOC_INF;
constexpr uint32_t FLAG_FIRST = 1u, FLAG_LAST = 2u;

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t addr = smem_u32(bar);
    uint32_t ok;
#ifdef BM25X_WATCHDOG
    uint32_t spins = 0;
#endif
    do {
#ifdef BM25X_WATCHDOG
        if (++spins > (1u << 26)) __trap();  // debug builds: turn a pipeline deadlock into a launch failure
#endif
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok) __nanosleep(40);  // do not burn issue slots of the merge warps while waiting
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}


template <int M_>
struct KCfg {
    static constexpr int M = M_;                        // max live terms per query in this class
    static constexpr int W = 8;                         // merge warps
    static constexpr int T = W * 32;                    // merge threads
    static constexpr int CB = (M_ <= 8) ? 16 : 32;      // 128-posting blocks per stage (>= M)
    static constexpr int STAGES = 3;
    static constexpr int QC = 512;                      // cold path: candidate queue entries per drain
    static constexpr int PC = 2048;                     // pool capacity (power of two >= BM25X_MAX_K + QC)
    static constexpr int QCW = 32;                      // hot path: per-warp candidate / possible-duplicate lists
    static constexpr int LOG_SW = (M_ <= 8) ? 11 : 12;  // per-warp tag map slots = 2^LOG_SW bytes
    static constexpr int STAGE_POSTINGS = CB * (int)BM25X_BLOCK;
    static constexpr int THREADS = T + 64;              // + producer warp + splitter warp
    static constexpr int MIN_CTAS = (M_ <= 8) ? 2 : 1;
};

template <int M>
struct Hdr {
    int qid;  // < 0: end of work
    uint32_t flags, lo, hi, m;
    uint32_t run_off[M], run_len[M];  // in postings, inside the stage
    float s0f[M];
    double s0d[M];
};

struct Ctrl {
    int qn, stall, pool_n, thr_valid, ovf, prefer_cold;
    int cnt[3];  // hot path: candidates appended by chunk n are counted in cnt[n % 3]
    float Flo;
    uint32_t dk, tie_sig;
    double Sk;
};

// Signature of a single-term document: (run, tf, fieldnorm).  Two documents with the same signature have bit-identical
// exact scores, so "same signature as the current k-th entry and a larger doc id" can be rejected without arithmetic.
constexpr uint32_t SIG_NONE = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t make_sig(uint32_t j, uint32_t w) {
    return (w >> 27) ? SIG_NONE : ((j << 27) | w);  // tf >= 2^19 does not fit beside the 5-bit run index
}

template <class C>
struct Smem {
    static constexpr size_t stage_bytes = (size_t)C::STAGE_POSTINGS * sizeof(Posting);
    static constexpr size_t off_stage = 0;
    static constexpr size_t off_map = off_stage + stage_bytes * C::STAGES;  // W private tag maps; cold path: bounds
    static constexpr size_t map_bytes = ((size_t)1 << C::LOG_SW) * C::W;
    static constexpr size_t off_hdr = off_map + map_bytes;
    static constexpr size_t hdr_bytes = (sizeof(Hdr<C::M>) + 15) & ~(size_t)15;
    static constexpr size_t off_bar = off_hdr + hdr_bytes * C::STAGES;       // full | ready | empty
    static constexpr size_t off_wb = off_bar + 24 * C::STAGES;               // sub-window bounds [STAGES][M][W+1] u16
    static constexpr size_t wb_bytes = (((size_t)C::M * (C::W + 1) * 2) + 15) & ~(size_t)15;
    static constexpr size_t off_pool_s = off_wb + wb_bytes * C::STAGES;
    static constexpr size_t off_pool_d = off_pool_s + (size_t)C::PC * 8;
    static constexpr size_t off_pool_g = off_pool_d + (size_t)C::PC * 4;     // tie signatures
    static constexpr size_t off_queue = off_pool_g + (size_t)C::PC * 4;      // cold: QC entries; hot: W x QCW
    static constexpr size_t queue_bytes = (size_t)(C::QC > C::W * C::QCW ? C::QC : C::W * C::QCW) * 4;
    static constexpr size_t off_dup = off_queue + queue_bytes;               // hot: W x QCW
    static constexpr size_t off_s1f = off_dup + (size_t)C::W * C::QCW * 4;
    static constexpr size_t off_ctrl = off_s1f + 256 * 4;
    static constexpr size_t total = off_ctrl + ((sizeof(Ctrl) + 15) & ~(size_t)15);
    static_assert((size_t)C::M * (C::T + 2) * 2 <= map_bytes, "cold-path bounds must fit in the tag maps");
    static_assert(C::PC >= 1024 + C::QC && (C::PC & (C::PC - 1)) == 0, "pool: power of two >= k + one drain");
    static_assert(C::PC >= 1024 + C::W * C::QCW, "pool must hold k + one chunk of hot-path candidates");
    static_assert(C::STAGE_POSTINGS <= 65536, "stage positions are 16-bit");
};

template <int T>
__device__ __forceinline__ void cbar() {  // barrier over the merge threads only (producer warp excluded)
    asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory");
}

__device__ __forceinline__ bool key_before(uint64_t ka, uint32_t da, uint64_t kb, uint32_t db) {
    return ka > kb || (ka == kb && da < db);  // score desc, doc asc (scores are > 0: raw f64 bits are monotone)
}

template <int LOG_S>
__device__ __forceinline__ uint32_t slot_of(uint32_t doc) {
    return (doc * 0x9E3779B1u) >> (32 - LOG_S);
}

// Cache::evaluate (bm25.rs:355-358) in f32, for the filter only.
__device__ __forceinline__ float score_f32(uint32_t w, float s0, const float *s1f) {
    float tff = (float)(w >> 8);
    float r;  // tf + s1 >= 1: no range guard needed around the approximate reciprocal (1 ulp)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(tff + s1f[w & 0xFFu]));
    return tff * s0 * r;
}
// Cache::evaluate in f64, bit-exact: (tf * s0) / (tf + s1[fieldnorm]).
__device__ __forceinline__ double score_f64(uint32_t w, double s0, const double *s1d) {
    double tfd = (double)(w >> 8);
    return __ddiv_rn(__dmul_rn(tfd, s0), __dadd_rn(tfd, s1d[w & 0xFFu]));
}

// Bitonic sort of the pool, best first.  n2 = power of two >= n.  Small pools are sorted by warp 0 alone.
template <int T>
__device__ void pool_sort(uint64_t *ks, uint32_t *ds, uint32_t *gs, int n, int n2, int tid) {
    for (int i = n + tid; i < n2; i += T) {
        ks[i] = 0;
        ds[i] = INF;
        gs[i] = SIG_NONE;
    }
    cbar<T>();
    const bool solo = n2 <= 64;  // 32 compare-exchanges per stage: one warp, __syncwarp between stages
    if (solo && tid >= 32) return;
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (n2 >> 1); i += (solo ? 32 : T)) {
                int a = 2 * i - (i & (stride - 1));
                int b = a + stride;
                uint64_t ka = ks[a], kb = ks[b];
                uint32_t da = ds[a], db = ds[b];
                bool desc = (a & size) == 0;
                bool sw = desc ? key_before(kb, db, ka, da) : key_before(ka, da, kb, db);
                if (sw) {
                    ks[a] = kb;
                    ks[b] = ka;
                    ds[a] = db;
                    ds[b] = da;
                    uint32_t ga = gs[a];
                    gs[a] = gs[b];
                    gs[b] = ga;
                }
            }
            if (solo) __syncwarp();
            else cbar<T>();
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <class C>
struct Bars {
    uint64_t *full, *ready, *empty;
    __device__ explicit Bars(uint8_t *smem) {
        full = (uint64_t *)(smem + Smem<C>::off_bar);
        ready = full + C::STAGES;
        empty = ready + C::STAGES;
    }
};

template <class C>
__device__ void producer(const SearchParams &p, uint8_t *smem, int lane) {
    constexpr int M = C::M;
    using S = Smem<C>;
    Bars<C> bars(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (;;) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(p.work_counter, 1);
        qi = __shfl_sync(0xFFFFFFFFu, qi, 0);
        if (qi >= (int)p.nq) break;
        const uint32_t qid = p.q_ids[qi];
        const uint32_t t0 = p.q_off[qi];
        const uint32_t m = p.q_off[qi + 1] - t0;  // 1..M
        uint32_t dfj = 0, nb = 0;
        uint64_t pbase = 0, bbase = 0;
        float s0f = 0.f;
        double s0d = 0.0;
        if (lane < (int)m) {
            uint32_t term = p.q_terms[t0 + lane];
            dfj = p.df[term];
            pbase = p.post_off[term];
            bbase = p.blk_off[term];
            nb = (dfj + BM25X_BLOCK - 1) / BM25X_BLOCK;
            s0f = p.s0f[term];
            s0d = p.s0d[term];
        }
        // block quota per term ∝ df: Σ quota <= CB
        uint64_t sumdf = dfj;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sumdf += __shfl_xor_sync(0xFFFFFFFFu, sumdf, o);
        const uint32_t quota_full = lane < (int)m ? 1u + (uint32_t)(((uint64_t)(C::CB - m) * dfj) / sumdf) : 0u;
        uint32_t ib = 0, lo = 0, chunk = 0;
        for (;;) {
            // warm-up: the first two chunks of a query are small (one block per term, the very first one cut to a
            // quarter of its window) so that the k-th-score threshold exists before the bulk of the postings arrives
            const uint32_t quota = lane < (int)m ? (chunk < 2 ? 1u : quota_full) : 0u;
            // window end: the smallest "first doc of the block just past my quota" over the terms
            uint32_t prop = INF;
            if (lane < (int)m && ib + quota < nb) prop = p.blk[bbase + ib + quota].x;
            uint32_t hi = __reduce_min_sync(0xFFFFFFFFu, prop);
            if (chunk == 0 && hi != INF) hi = lo + max(1u, (hi - lo) >> 2);
            uint32_t eb = ib, lastd = 0;
            if (lane < (int)m) {
                uint32_t lim = min(nb, ib + quota);
                for (uint32_t b = ib; b < lim; ++b) {
                    uint2 d = p.blk[bbase + b];
                    if (d.x < hi) {
                        eb = b + 1;
                        lastd = d.y;
                    }
                }
            }
            uint32_t len = 0;
            if (eb > ib) {
                uint32_t endp = min(eb * BM25X_BLOCK, dfj);
                len = (endp - ib * BM25X_BLOCK + 1u) & ~1u;  // whole 16-byte units; the odd tail is a pad slot
            }
            uint32_t incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            const uint32_t off = incl - len;
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            const bool last = hi == INF;

            mbar_wait(&bars.empty[stage], phase ^ 1u);
            Hdr<M> *h = (Hdr<M> *)(smem + S::off_hdr + S::hdr_bytes * stage);
            if (lane < M) {
                h->run_off[lane] = off;
                h->run_len[lane] = len;
                h->s0f[lane] = s0f;
                h->s0d[lane] = s0d;
            }
            if (lane == 0) {
                h->qid = (int)qid;
                h->flags = (chunk == 0 ? FLAG_FIRST : 0u) | (last ? FLAG_LAST : 0u);
                h->lo = lo;
                h->hi = min(hi, p.n_docs);
                h->m = m;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(&bars.full[stage], total * (uint32_t)sizeof(Posting));
            __syncwarp();
            if (len > 0) {
                tma_load_1d(smem + S::off_stage + S::stage_bytes * stage + (size_t)off * sizeof(Posting),
                            p.post + pbase + (uint64_t)ib * BM25X_BLOCK, len * (uint32_t)sizeof(Posting),
                            &bars.full[stage]);
            }
            if (eb > ib) ib = (lastd >= hi) ? eb - 1 : eb;  // keep a block that straddles the window end
            lo = hi;
            chunk++;
            if (++stage == C::STAGES) {
                stage = 0;
                phase ^= 1u;
            }
            if (last) break;
        }
    }
    mbar_wait(&bars.empty[stage], phase ^ 1u);
    if (lane == 0) {
        Hdr<M> *h = (Hdr<M> *)(smem + S::off_hdr + S::hdr_bytes * stage);
        h->qid = -1;
        mbar_arrive(&bars.full[stage]);
    }
}

// ---------------------------------------------------------------------------------------------
// Splitter warp: as soon as a chunk has landed, cut its doc window [lo, hi) into W equal sub-windows and find, for
// every run, the stage position of each boundary (lower_bound).  wb[j][b] = first posting of run j with doc >= x_b.
template <class C>
__device__ void splitter(uint8_t *smem, int lane) {
    constexpr int M = C::M;
    constexpr int W = C::W;
    using S = Smem<C>;
    Bars<C> bars(smem);
    int stage = 0;
    uint32_t phase = 0;
    for (;;) {
        mbar_wait(&bars.full[stage], phase);
        const Hdr<M> *h = (const Hdr<M> *)(smem + S::off_hdr + S::hdr_bytes * stage);
        const bool end = h->qid < 0;
        if (!end) {
            const Posting *st = (const Posting *)(smem + S::off_stage + S::stage_bytes * stage);
            uint16_t *wb = (uint16_t *)(smem + S::off_wb + S::wb_bytes * stage);
            const uint32_t lo = h->lo, hi = h->hi, m = h->m;
            const uint64_t span = (uint64_t)hi - lo;
            for (uint32_t it = lane; it < m * (W + 1); it += 32) {
                const uint32_t j = it / (W + 1), b = it - j * (W + 1);
                const uint32_t x = b == W ? hi : lo + (uint32_t)((span * b) / W);
                const uint32_t a = h->run_off[j];
                uint32_t l = 0, r = h->run_len[j];
                while (l < r) {
                    uint32_t mid = (l + r) >> 1;
                    if (st[a + mid].doc < x) l = mid + 1;
                    else r = mid;
                }
                wb[j * (W + 1) + b] = (uint16_t)(a + l);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.ready[stage]);
        if (end) break;
        if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1u;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Shared by both paths: everything a merge thread needs to know about the current chunk.
template <class C>
struct Chunk {
    const Hdr<C::M> *h;
    const Posting *st;
    uint32_t m, lo, span;
    uint8_t *maps;  // W private tag maps (hot) / bucket bounds (cold)
    uint64_t *pool_s;
    uint32_t *pool_d, *pool_g;
    uint32_t *queue;
    const float *s1f;
    volatile Ctrl *ctrl;
};

// lower_bound of `doc` in stage positions [a, e); returns the posting word or 0 when absent
__device__ __forceinline__ uint32_t find_in(const Posting *st, uint32_t a, uint32_t e, uint32_t doc) {
    uint32_t l = a, r = e;
    while (l < r) {
        uint32_t mid = (l + r) >> 1;
        if (st[mid].doc < doc) l = mid + 1;
        else r = mid;
    }
    if (l < e) {
        Posting v = st[l];
        if (v.doc == doc) return v.w;
    }
    return 0u;
}

// The threshold filter on the f32 score F of a complete document: F below Sk·(1-2^-18) cannot reach the top-k
// (f32 error bound); a single-term document with the signature of the current k-th entry has exactly
// the k-th score and enters only with a smaller doc id.  Everything else is re-scored exactly.
struct Filter {
    bool tv;
    float Flo;
    double Sk;
    uint32_t dk, tie_sig, tie_dk;  // tie_dk = INF when there is no usable tie signature
};
__device__ __forceinline__ bool filter_pass(const Filter &f, float F, uint32_t sig, uint32_t doc) {
    return F >= f.Flo && !(sig == f.tie_sig && doc > f.tie_dk);
}

template <class C>
__device__ __forceinline__ Filter load_filter(const Chunk<C> &c) {
    Filter f;
    f.tv = c.ctrl->thr_valid != 0;
    f.Flo = f.tv ? c.ctrl->Flo : -1.f;  // scores are > 0: -1 lets everything through
    f.tie_sig = c.ctrl->tie_sig;
    f.Sk = c.ctrl->Sk;
    f.dk = c.ctrl->dk;
    f.tie_dk = (f.tv && f.tie_sig != SIG_NONE) ? f.dk : INF;
    return f;
}

template <class C>
__device__ __forceinline__ void clear_maps(const Chunk<C> &c, int tid) {
    uint4 *mp = (uint4 *)c.maps;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < (int)(Smem<C>::map_bytes / 16); i += C::T) mp[i] = z;
}

// Pool upkeep after a CTA barrier: cut the pool back to k when it is getting full (or for the final output) and
// refresh the k-th-score threshold.  All merge threads call; contains barriers only when it sorts.
// `pn` = current pool size, identical in every thread (uniform register); returns the new size.
template <class C>
__device__ int upkeep(const Chunk<C> &c, const SearchParams &p, int pn, bool fin, int tid) {
    constexpr int T = C::T;
    const double kEps = 1.0 / 262144.0;  // 2^-18 > f32 error bound of the filter score
    const uint32_t k = p.k;
    const int slack = (int)k > 22 ? (int)k : 22;
    const int room = C::QC > C::W * C::QCW ? C::QC : C::W * C::QCW;
    const bool need = pn > 0 && (fin || pn > C::PC - room || pn >= (int)k + slack);
    if (!need) return pn;
    int n2 = 2;
    while (n2 < pn) n2 <<= 1;
    pool_sort<T>(c.pool_s, c.pool_d, c.pool_g, pn, n2, tid);
    cbar<T>();
    const int nn = pn < (int)k ? pn : (int)k;
    if (tid == 0) {
        if (nn == (int)k) {
            double sk = __longlong_as_double((long long)c.pool_s[k - 1]);
            c.ctrl->Sk = sk;
            c.ctrl->dk = c.pool_d[k - 1];
            c.ctrl->tie_sig = c.pool_g[k - 1];
            c.ctrl->Flo = __double2float_rd(sk * (1.0 - kEps));
            c.ctrl->thr_valid = 1;
        }
    }
    cbar<T>();
    return nn;
}

// Exact f64 score of `doc` over the runs' stage ranges [ra[j], re[j]) and its posting count.
// Reference order: Cache::evaluate per term (bm25.rs:355-358), terms ascending.
template <class C, class RA, class RE>
__device__ __forceinline__ double exact_score(const Chunk<C> &c, const SearchParams &p, uint32_t doc, RA ra, RE re,
                                              uint32_t &cnt, uint32_t &sig) {
    double Sx = 0.0;
    cnt = 0;
    for (uint32_t j = 0; j < c.m; ++j) {
        uint32_t w = find_in(c.st, ra(j), re(j), doc);
        if (w) {
            Sx = __dadd_rn(Sx, score_f64(w, c.h->s0d[j], p.s1d));
            cnt++;
            sig = make_sig(j, w);
        }
    }
    return Sx;
}

// ---------------------------------------------------------------------------------------------
// Cold path (kernel v1): equal-width doc-id buckets, one per thread, posting-at-a-time m-way merge with the heads in
// registers.  Handles any overlap density and candidate-queue overflow (rounds).  Uses the tag-map memory for the
// bucket bounds; leaves the maps zeroed.  All merge threads call, after a CTA barrier.
template <class C>
__device__ __noinline__ int cold_chunk(const Chunk<C> &c, const SearchParams &p, int pn, bool last, int tid) {
    constexpr int M = C::M;
    constexpr int T = C::T;
    uint16_t *bounds = (uint16_t *)c.maps;
    const Hdr<M> *h = c.h;
    const Posting *st = c.st;
    if (tid == 0) {
        c.ctrl->qn = 0;
        c.ctrl->ovf = 0;
        c.ctrl->stall = 0;
        c.ctrl->pool_n = pn;  // the cold path appends through the shared counter, between CTA barriers
    }
    {
        const uint32_t lo = c.lo;
        const uint32_t hi = c.lo + c.span;
        const uint64_t mult = ((uint64_t)T << 32) / (uint64_t)c.span;
        auto keyb = [&](uint32_t d) -> uint32_t {
            if (d < lo) return 0u;
            if (d >= hi) return (uint32_t)T + 1u;
            return 1u + (uint32_t)(((uint64_t)(d - lo) * mult) >> 32);
        };
#pragma unroll 1
        for (int j = 0; j < M; ++j) {
            const uint32_t off = h->run_off[j], len = h->run_len[j];
            uint16_t *B = bounds + j * (T + 2);
            for (uint32_t i = tid; i <= len; i += T) {
                uint32_t kc = i < len ? keyb(st[off + i].doc) : (uint32_t)T + 1u;
                uint32_t kp = i == 0 ? 0u : keyb(st[off + i - 1].doc);
#pragma unroll 1
                for (uint32_t cc = kp + 1; cc <= kc; ++cc) B[cc] = (uint16_t)(off + i);
            }
        }
    }
    cbar<T>();
    uint32_t hd[M], hw[M], pp[M], pe[M];
    float s0r[M];
#pragma unroll
    for (int j = 0; j < M; ++j) {
        pp[j] = bounds[j * (T + 2) + tid + 1];
        pe[j] = bounds[j * (T + 2) + tid + 2];
        s0r[j] = h->s0f[j];
        hd[j] = INF;
        hw[j] = 0;
        if (pp[j] < pe[j]) {
            Posting v = st[pp[j]];
            hd[j] = v.doc;
            hw[j] = v.w;
        }
    }
    cbar<T>();  // everyone holds its bounds in registers: the map memory can be zeroed again
    clear_maps(c, tid);
    uint32_t cur = INF, cnt = 0, lj = 0, lw = 0;
    float F = 0.f;
    bool done = false;
    for (;;) {
        Filter f = load_filter(c);
        while (!done) {
            uint32_t dmin = hd[0], wm = hw[0];
            float s0m = s0r[0];
            int jm = 0;
#pragma unroll
            for (int j = 1; j < M; ++j) {
                bool lt = hd[j] < dmin;  // strict: equal docs are consumed in ascending term order
                dmin = lt ? hd[j] : dmin;
                wm = lt ? hw[j] : wm;
                s0m = lt ? s0r[j] : s0m;
                jm = lt ? j : jm;
            }
            if (dmin != cur) {
                if (cur != INF && filter_pass(f, F, cnt == 1 ? make_sig(lj, lw) : SIG_NONE, cur)) {
                    int idx = atomicAdd((int *)&c.ctrl->qn, 1);
                    if (idx < C::QC) {
                        c.queue[idx] = cur;
                    } else {
                        c.ctrl->stall = 1;  // queue full: retry this document after the drain
                        break;
                    }
                }
                cur = dmin;
                F = 0.f;
                cnt = 0;
            }
            if (dmin == INF) {
                done = true;
                break;
            }
            F += score_f32(wm, s0m, c.s1f);
            cnt++;
            lj = (uint32_t)jm;
            lw = wm;
            uint32_t np = 0, ne = 0;
#pragma unroll
            for (int j = 0; j < M; ++j) {
                if (j == jm) {
                    pp[j] += 1;
                    np = pp[j];
                    ne = pe[j];
                }
            }
            uint32_t nd = INF, nw = 0;
            if (np < ne) {
                Posting v = st[np];
                nd = v.doc;
                nw = v.w;
            }
#pragma unroll
            for (int j = 0; j < M; ++j) {
                if (j == jm) {
                    hd[j] = nd;
                    hw[j] = nw;
                }
            }
        }
        cbar<T>();
        // drain: exact f64 re-score of the queued documents over the whole runs
        const int stalled = c.ctrl->stall;
        const int nqueue = min(c.ctrl->qn, C::QC);
        for (int e = tid; e < nqueue; e += T) {
            const uint32_t doc = c.queue[e];
            if (p.allow && !((p.allow[doc >> 3] >> (doc & 7u)) & 1u)) continue;  // filter(payload), search.rs:230
            uint32_t n = 0, sig = SIG_NONE;
            double Sx = exact_score(
                c, p, doc, [&](uint32_t j) { return h->run_off[j]; },
                [&](uint32_t j) { return h->run_off[j] + h->run_len[j]; }, n, sig);
            if (!f.tv || Sx > f.Sk || (Sx == f.Sk && doc < f.dk)) {
                int idx = atomicAdd((int *)&c.ctrl->pool_n, 1);
                c.pool_s[idx] = (uint64_t)__double_as_longlong(Sx);
                c.pool_d[idx] = doc;
                c.pool_g[idx] = n == 1 ? sig : SIG_NONE;
            }
        }
        cbar<T>();
        pn = c.ctrl->pool_n;  // stable: nobody appends until the next round
        cbar<T>();
        if (tid == 0) {
            c.ctrl->qn = 0;
            c.ctrl->stall = 0;
        }
        const int nn = upkeep(c, p, pn, !stalled && last, tid);
        if (tid == 0) c.ctrl->pool_n = nn;
        pn = nn;
        cbar<T>();
        if (!stalled) break;
    }
    return pn;
}

// ---------------------------------------------------------------------------------------------
// Hot path of one merge warp on its doc sub-window of the chunk.  Everything is warp-private except the pool append.
template <class C>
__device__ __forceinline__ void hot_warp(const Chunk<C> &c, const SearchParams &p, const uint16_t *wb, int cw, int lane,
                                         uint8_t *map, uint32_t *wq, uint32_t *wd, int pn, volatile int *append_counter) {
    constexpr int W = C::W;
    const Posting *st = c.st;
    const Hdr<C::M> *h = c.h;
    const uint32_t m = c.m;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // my sub-run of run `lane`
    uint32_t my_a = 0, my_e = 0;
    if (lane < (int)m) {
        my_a = wb[lane * (W + 1) + cw];
        my_e = wb[lane * (W + 1) + cw + 1];
    }
    Filter f = load_filter(c);
    // ---- A: mark ----
    for (uint32_t j = 0; j < m; ++j) {
        const uint32_t a = __shfl_sync(0xFFFFFFFFu, my_a, j), e = __shfl_sync(0xFFFFFFFFu, my_e, j);
        const uint8_t tagv = (uint8_t)(j + 1);
        for (uint32_t i = a + lane; i < e; i += 32) map[slot_of<C::LOG_SW>(st[i].doc)] = tagv;
    }
    __syncwarp();
    // ---- B: test; score + filter the singles; list the possible duplicates ----
    uint32_t nd = 0, nc = 0;  // warp-uniform list lengths
    for (uint32_t j = 0; j < m; ++j) {
        const uint32_t a = __shfl_sync(0xFFFFFFFFu, my_a, j), e = __shfl_sync(0xFFFFFFFFu, my_e, j);
        const float s0 = h->s0f[j];
        const uint8_t tagv = (uint8_t)(j + 1);
        for (uint32_t base = a; base < e; base += 32) {
            const uint32_t i = base + lane;
            const bool valid = i < e;
            Posting v = st[valid ? i : a];
            const bool dup = valid && map[slot_of<C::LOG_SW>(v.doc)] != tagv;
            const bool cand = valid && !dup && filter_pass(f, score_f32(v.w, s0, c.s1f), make_sig(j, v.w), v.doc);
            const uint32_t md = __ballot_sync(0xFFFFFFFFu, dup);
            if (md) {
                const uint32_t pos = nd + __popc(md & lt_mask);
                if (dup && pos < (uint32_t)C::QCW) wd[pos] = (j << 16) | i;
                nd += __popc(md);
            }
            const uint32_t mc = __ballot_sync(0xFFFFFFFFu, cand);
            if (mc) {
                const uint32_t pos = nc + __popc(mc & lt_mask);
                if (cand && pos < (uint32_t)C::QCW) wq[pos] = (j << 16) | i;
                nc += __popc(mc);
            }
        }
    }
    __syncwarp();
    if (nd > (uint32_t)C::QCW) {  // dense overlap: let the whole chunk go through the merge path
        if (lane == 0) {
            c.ctrl->ovf = 1;
            c.ctrl->prefer_cold = 1;
        }
        return;
    }
    // ---- C: resolve the possible duplicates; exactly one emitter per document ----
    if (nd) {
        const bool has = lane < (int)nd;
        const uint32_t ent = has ? wd[lane] : 0u;
        const uint32_t j = ent >> 16;
        const Posting v = st[ent & 0xFFFFu];
        const uint32_t winner = (uint32_t)map[slot_of<C::LOG_SW>(v.doc)] - 1u;
        float F = 0.f;
        uint32_t cnt = 0;
        bool owner = has;
        for (uint32_t jj = 0; jj < m; ++jj) {
            const uint32_t a = __shfl_sync(0xFFFFFFFFu, my_a, jj), e = __shfl_sync(0xFFFFFFFFu, my_e, jj);
            if (!owner) continue;
            const uint32_t w = jj == j ? v.w : find_in(st, a, e, v.doc);
            if (!w) continue;
            if (jj < j && jj != winner) {  // a lower run also detected this document: it emits
                owner = false;
                continue;
            }
            F += score_f32(w, h->s0f[jj], c.s1f);
            cnt++;
        }
        const bool cand = owner && filter_pass(f, F, cnt == 1 ? make_sig(j, v.w) : SIG_NONE, v.doc);
        const uint32_t mc = __ballot_sync(0xFFFFFFFFu, cand);
        if (mc) {
            const uint32_t pos = nc + __popc(mc & lt_mask);
            if (cand && pos < (uint32_t)C::QCW) wq[pos] = ent;
            nc += __popc(mc);
        }
        __syncwarp();
    }
    if (nc > (uint32_t)C::QCW) {
        if (lane == 0) c.ctrl->ovf = 1;
        return;
    }
    // ---- D: exact f64 re-score of the survivors → pool ----
    if (nc) {
        const bool has = lane < (int)nc;
        const uint32_t ent = has ? wq[lane] : 0u;
        const uint32_t doc = st[ent & 0xFFFFu].doc;
        double Sx = 0.0;
        uint32_t cnt = 0, sig = SIG_NONE;
        for (uint32_t jj = 0; jj < m; ++jj) {
            const uint32_t a = __shfl_sync(0xFFFFFFFFu, my_a, jj), e = __shfl_sync(0xFFFFFFFFu, my_e, jj);
            if (!has) continue;
            const uint32_t w = find_in(st, a, e, doc);
            if (w) {
                Sx = __dadd_rn(Sx, score_f64(w, h->s0d[jj], p.s1d));
                cnt++;
                sig = make_sig(jj, w);
            }
        }
        bool keep = has;
        if (keep && p.allow && !((p.allow[doc >> 3] >> (doc & 7u)) & 1u)) keep = false;  // filter(payload), search.rs:230
        // a posting whose tag won its slot although other runs hold the document: its detecting twin carries it
        if (keep && cnt > 1 && map[slot_of<C::LOG_SW>(doc)] == (uint8_t)((ent >> 16) + 1u)) keep = false;
        keep = keep && (!f.tv || Sx > f.Sk || (Sx == f.Sk && doc < f.dk));
        const uint32_t mk = __ballot_sync(0xFFFFFFFFu, keep);
        if (mk) {  // pool slots [pn + base, ...): pn is the (uniform) pool size at the start of the chunk
            int base = 0;
            if (lane == 0) base = atomicAdd((int *)append_counter, __popc(mk));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (keep) {
                const int idx = pn + base + __popc(mk & lt_mask);
                if (idx < C::PC) {
                    c.pool_s[idx] = (uint64_t)__double_as_longlong(Sx);
                    c.pool_d[idx] = doc;
                    c.pool_g[idx] = cnt == 1 ? sig : SIG_NONE;
                } else {
                    c.ctrl->ovf = 1;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <class C>
__device__ void consumer(const SearchParams &p, uint8_t *smem, int tid) {
    constexpr int M = C::M;
    constexpr int T = C::T;
    using S = Smem<C>;
    const int lane = tid & 31, cw = tid >> 5;
    Bars<C> bars(smem);
    Chunk<C> c;
    c.maps = smem + S::off_map;
    c.pool_s = (uint64_t *)(smem + S::off_pool_s);
    c.pool_d = (uint32_t *)(smem + S::off_pool_d);
    c.pool_g = (uint32_t *)(smem + S::off_pool_g);
    c.queue = (uint32_t *)(smem + S::off_queue);
    c.s1f = (const float *)(smem + S::off_s1f);
    c.ctrl = (volatile Ctrl *)(smem + S::off_ctrl);
    uint8_t *my_map = c.maps + ((size_t)cw << C::LOG_SW);
    uint32_t *my_q = c.queue + cw * C::QCW;
    uint32_t *my_d = (uint32_t *)(smem + S::off_dup) + cw * C::QCW;
    const uint32_t k = p.k;

    int stage = 0;
    uint32_t phase = 0;
    int pn = 0;         // pool size — the same value in every merge thread
    uint32_t seq = 0;   // chunk sequence number of this CTA (selects the append counter)
    for (;;) {
        mbar_wait(&bars.ready[stage], phase);
        const Hdr<M> *h = (const Hdr<M> *)(smem + S::off_hdr + S::hdr_bytes * stage);
        const int qid = h->qid;
        if (qid < 0) break;
        const bool last = (h->flags & FLAG_LAST) != 0;
        c.h = h;
        c.st = (const Posting *)(smem + S::off_stage + S::stage_bytes * stage);
        c.m = h->m;
        c.lo = h->lo;
        c.span = h->hi - h->lo;
        const uint32_t slot = seq % 3u;
        const bool go_cold = c.ctrl->prefer_cold != 0;
        if (!go_cold) {
            hot_warp<C>(c, p, (const uint16_t *)(smem + S::off_wb + S::wb_bytes * stage), cw, lane, my_map, my_q, my_d, pn,
                        &c.ctrl->cnt[slot]);
            __syncwarp();
            uint4 *mp = (uint4 *)my_map;  // zero my tag map for the next chunk
            for (int i = lane; i < (1 << C::LOG_SW) / 16; i += 32) mp[i] = make_uint4(0, 0, 0, 0);
        }
        cbar<T>();  // the one CTA barrier of the hot path: all sub-windows of the chunk are done
        // everything read below was written before the barrier and is not written again before the next one
        const bool cold = go_cold || c.ctrl->ovf != 0;
        const int added = c.ctrl->cnt[slot];
        if (tid == 0) c.ctrl->cnt[(seq + 2u) % 3u] = 0;  // the counter of the previous chunk: everyone has read it
        if (cold) {
            cbar<T>();  // everyone has read ovf; the hot attempt's appends are simply not counted (pn unchanged)
            pn = cold_chunk<C>(c, p, pn, last, tid);
        } else {
            pn = upkeep(c, p, pn + added, last, tid);
        }
        if (last) {  // Results::into_sorted_vec (search.rs:281): the pool is sorted, best first
            const int n = pn;
            const size_t base = (size_t)qid * k;
            for (int i = tid; i < (int)k; i += T) {
                uint32_t d = INF;
                double sc = 0.0;
                if (i < n) {
                    d = c.pool_d[i];
                    sc = __longlong_as_double((long long)c.pool_s[i]);
                }
                p.out_doc[base + i] = d;
                p.out_score[base + i] = (float)sc;
                if (p.out_score64) p.out_score64[base + i] = sc;
                if (p.out_payload) {
                    uint16_t a = 0, b = 0, cc = 0;
                    if (i < n) {
                        a = p.payload[(size_t)d * 3 + 0];
                        b = p.payload[(size_t)d * 3 + 1];
                        cc = p.payload[(size_t)d * 3 + 2];
                    }
                    p.out_payload[(base + i) * 3 + 0] = a;
                    p.out_payload[(base + i) * 3 + 1] = b;
                    p.out_payload[(base + i) * 3 + 2] = cc;
                }
            }
            cbar<T>();  // the pool has been read out
            if (tid == 0) {  // reset for the next query
                p.out_n[qid] = (uint32_t)n;
                c.ctrl->thr_valid = 0;
                c.ctrl->prefer_cold = 0;
            }
            pn = 0;
            cbar<T>();
        }
        seq++;
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.empty[stage]);
        if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1u;
        }
    }
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::MIN_CTAS) k_search(const __grid_constant__ SearchParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    using S = Smem<C>;
    const int tid = threadIdx.x;
    if (tid == 0) {
        Bars<C> bars(smem);
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&bars.full[s], 1);
            mbar_init(&bars.ready[s], 1);
            mbar_init(&bars.empty[s], C::W);
        }
        mbar_fence_init();
        Ctrl *ctrl = (Ctrl *)(smem + S::off_ctrl);
        ctrl->qn = 0;
        ctrl->stall = 0;
        ctrl->pool_n = 0;
        ctrl->thr_valid = 0;
        ctrl->ovf = 0;
        ctrl->prefer_cold = 0;
        ctrl->cnt[0] = ctrl->cnt[1] = ctrl->cnt[2] = 0;
        ctrl->tie_sig = SIG_NONE;
    }
    for (int i = tid; i < 256; i += C::THREADS) ((float *)(smem + S::off_s1f))[i] = p.s1f[i];
    for (int i = tid; i < (int)(S::map_bytes / 16); i += C::THREADS) ((uint4 *)(smem + S::off_map))[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (tid < C::T) consumer<C>(p, smem, tid);
    else if (tid < C::T + 32) producer<C>(p, smem, tid - C::T);
    else splitter<C>(smem, tid - C::T - 32);
}

}  // namespace
