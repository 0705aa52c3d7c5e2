// bm25x_search.cu — batched BM25 top-k over the HBM-resident index (sm_100a).
//
// Replaces bm25::search (crates/bm25/src/search.rs:28-282): instead of one query walking cursors
// over 8 KiB pages with Block-max WAND, a persistent grid streams every query's posting lists
// through shared memory with TMA bulk copies and merges them there.
//
//   producer warp  : per query, cuts the doc-id space into chunks whose postings fit one smem stage
//                    (quota of 128-posting blocks per term ∝ df, boundaries from the per-block
//                    (first doc, last doc) table = SummaryTuple.{min,max}_document_id), and issues one
//                    cp.async.bulk per term per chunk, completion on an mbarrier (STAGES-deep ring).
//   merge threads  : phase 0 buckets the chunk's postings by doc id (T equal-width buckets; one pass,
//                    no search); phase A: each thread m-way merges its bucket (posting at a time,
//                    heads in registers), f32 score per doc = Σ s0·tf/(tf+s1[fn]) — the Cache::evaluate
//                    formula (bm25.rs:355-358) in f32 — and filters against the current k-th score;
//                    phase B: the few survivors are re-scored in f64 with the reference's exact
//                    operation order and appended to a pool; the pool is cut back to k by a bitonic
//                    sort when it fills.  Final order: score desc, doc id asc.
//
// Exactness: the f32 filter only ever *rejects* documents whose f32 score is below
// Sk·(1-2^-18) where Sk is the exact f64 k-th best so far; the f32 error bound is < 2^-18
// relative, so no document of the true top-k is ever rejected; everything that
// survives is ranked by its exact f64 score.
#include <algorithm>
#include <vector>

#include "bm25x_common.h"

namespace {

constexpr uint32_t INF = BM25X_DOC_INF;
constexpr uint32_t FLAG_FIRST = 1u, FLAG_LAST = 2u;

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk async copy (TMA), as in the Blackwell guide §15.
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
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
struct SearchParams {
    const Posting *post;
    const uint64_t *post_off;
    const uint32_t *df;
    const uint64_t *blk_off;
    const uint2 *blk;
    const float *s0f;
    const double *s0d;
    const double *s1d;
    const float *s1f;
    const uint16_t *payload;
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

template <int M_>
struct KCfg {
    static constexpr int M = M_;                      // max live terms per query in this class
    static constexpr int T = 256;                     // merge threads (= doc-id buckets per chunk)
    static constexpr int CB = (M_ <= 8) ? 24 : 32;    // 128-posting blocks per stage (>= M)
    static constexpr int STAGES = 3;
    static constexpr int QC = 1024;                   // candidate queue entries per round
    static constexpr int PC = 2048;                   // pool capacity (>= BM25X_MAX_K + QC)
    static constexpr int STAGE_POSTINGS = CB * (int)BM25X_BLOCK;
    static constexpr int THREADS = T + 32;
    static constexpr int MIN_CTAS = (M_ <= 4) ? 2 : 1;  // register budget: 2 CTAs/SM for the small classes
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
    int qn, stall, pool_n, thr_valid;
    float Flo, Fhi;
    double Sk;
    uint32_t dk;
};

template <class C>
struct Smem {
    static constexpr size_t stage_bytes = (size_t)C::STAGE_POSTINGS * sizeof(Posting);
    static constexpr size_t off_stage = 0;
    static constexpr size_t off_hdr = off_stage + stage_bytes * C::STAGES;
    static constexpr size_t hdr_bytes = (sizeof(Hdr<C::M>) + 15) & ~(size_t)15;
    static constexpr size_t off_bar = off_hdr + hdr_bytes * C::STAGES;
    static constexpr size_t off_bounds = off_bar + 16 * C::STAGES;
    static constexpr size_t bounds_bytes = (((size_t)C::M * (C::T + 2) * 2) + 15) & ~(size_t)15;
    static constexpr size_t off_pool_s = off_bounds + bounds_bytes;
    static constexpr size_t off_pool_d = off_pool_s + (size_t)C::PC * 8;
    static constexpr size_t off_queue = off_pool_d + (size_t)C::PC * 4;
    static constexpr size_t off_s1f = off_queue + (size_t)C::QC * 4;
    static constexpr size_t off_ctrl = off_s1f + 256 * 4;
    static constexpr size_t total = off_ctrl + ((sizeof(Ctrl) + 15) & ~(size_t)15);
};

template <int T>
__device__ __forceinline__ void cbar() {  // barrier over the merge threads only (producer warp excluded)
    asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory");
}

__device__ __forceinline__ bool key_before(uint64_t ka, uint32_t da, uint64_t kb, uint32_t db) {
    return ka > kb || (ka == kb && da < db);  // score desc, doc asc (scores are > 0: raw f64 bits are monotone)
}

// Bitonic sort of the pool, best first.  n2 = power of two >= n.
template <int T>
__device__ void pool_sort(uint64_t *ks, uint32_t *ds, int n, int n2, int tid) {
    for (int i = n + tid; i < n2; i += T) {
        ks[i] = 0;
        ds[i] = INF;
    }
    cbar<T>();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (n2 >> 1); i += T) {
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
                }
            }
            cbar<T>();
        }
    }
}

// ---------------------------------------------------------------------------------------------
template <class C>
__device__ void producer(const SearchParams &p, uint8_t *smem, int lane) {
    constexpr int M = C::M;
    using S = Smem<C>;
    uint64_t *full = (uint64_t *)(smem + S::off_bar);
    uint64_t *empty = full + C::STAGES;
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
        const uint32_t quota = lane < (int)m ? 1u + (uint32_t)(((uint64_t)(C::CB - m) * dfj) / sumdf) : 0u;
        uint32_t ib = 0, lo = 0;
        bool first = true;
        for (;;) {
            // window end: the smallest "first doc of the block just past my quota" over the terms
            uint32_t prop = INF;
            if (lane < (int)m && ib + quota < nb) prop = p.blk[bbase + ib + quota].x;
            const uint32_t hi = __reduce_min_sync(0xFFFFFFFFu, prop);
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

            mbar_wait(&empty[stage], phase ^ 1u);
            Hdr<M> *h = (Hdr<M> *)(smem + S::off_hdr + S::hdr_bytes * stage);
            if (lane < M) {
                h->run_off[lane] = off;
                h->run_len[lane] = len;
                h->s0f[lane] = s0f;
                h->s0d[lane] = s0d;
            }
            if (lane == 0) {
                h->qid = (int)qid;
                h->flags = (first ? FLAG_FIRST : 0u) | (last ? FLAG_LAST : 0u);
                h->lo = lo;
                h->hi = min(hi, p.n_docs);
                h->m = m;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_expect_tx(&full[stage], total * (uint32_t)sizeof(Posting));
            __syncwarp();
            if (len > 0) {
                tma_load_1d(smem + S::off_stage + S::stage_bytes * stage + (size_t)off * sizeof(Posting),
                            p.post + pbase + (uint64_t)ib * BM25X_BLOCK, len * (uint32_t)sizeof(Posting), &full[stage]);
            }
            if (eb > ib) ib = (lastd >= hi) ? eb - 1 : eb;  // keep a block that straddles the window end
            lo = hi;
            first = false;
            if (++stage == C::STAGES) {
                stage = 0;
                phase ^= 1u;
            }
            if (last) break;
        }
    }
    mbar_wait(&empty[stage], phase ^ 1u);
    if (lane == 0) {
        Hdr<M> *h = (Hdr<M> *)(smem + S::off_hdr + S::hdr_bytes * stage);
        h->qid = -1;
        mbar_arrive(&full[stage]);
    }
}

// ---------------------------------------------------------------------------------------------
template <class C>
__device__ void consumer(const SearchParams &p, uint8_t *smem, int tid) {
    constexpr int M = C::M;
    constexpr int T = C::T;
    using S = Smem<C>;
    const int lane = tid & 31;
    uint64_t *full = (uint64_t *)(smem + S::off_bar);
    uint64_t *empty = full + C::STAGES;
    uint16_t *bounds = (uint16_t *)(smem + S::off_bounds);
    uint64_t *pool_s = (uint64_t *)(smem + S::off_pool_s);
    uint32_t *pool_d = (uint32_t *)(smem + S::off_pool_d);
    uint32_t *q_doc = (uint32_t *)(smem + S::off_queue);
    const float *s1f = (const float *)(smem + S::off_s1f);
    volatile Ctrl *ctrl = (volatile Ctrl *)(smem + S::off_ctrl);
    const uint32_t k = p.k;
    const double kEps = 1.0 / 262144.0;  // 2^-18 > f32 error bound of the filter score

    int stage = 0;
    uint32_t phase = 0;
    for (;;) {
        mbar_wait(&full[stage], phase);
        const Hdr<M> *h = (const Hdr<M> *)(smem + S::off_hdr + S::hdr_bytes * stage);
        const int qid = h->qid;
        if (qid < 0) break;
        const uint32_t flags = h->flags;
        const uint32_t m = h->m;
        const Posting *st = (const Posting *)(smem + S::off_stage + S::stage_bytes * stage);
        if (flags & FLAG_FIRST) {
            if (tid == 0) {
                ctrl->pool_n = 0;
                ctrl->thr_valid = 0;
                ctrl->qn = 0;
                ctrl->stall = 0;
            }
        }
        // ---- phase 0: bucket boundaries of every run, one pass over the postings ----
        {
            const uint32_t lo = h->lo, hi = h->hi;
            const uint64_t span = (uint64_t)hi - lo;  // >= 1
            const uint64_t mult = ((uint64_t)T << 32) / span;
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
                    for (uint32_t c = kp + 1; c <= kc; ++c) B[c] = (uint16_t)(off + i);
                }
            }
        }
        cbar<T>();
        // ---- merge state: heads of my bucket's sub-runs in registers ----
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
        uint32_t cur = INF, cnt = 0, lj = 0, lw = 0;
        float F = 0.f;
        bool done = false;
        for (;;) {  // rounds: phase A (merge + filter) → phase B (exact re-score) → pool upkeep
            const bool tv = ctrl->thr_valid != 0;
            const float Flo = ctrl->Flo, Fhi = ctrl->Fhi;
            const double Sk = ctrl->Sk;
            const uint32_t dk = ctrl->dk;
            uint32_t cj = INF, cw = 0;  // single-term signature known to score exactly Sk
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
                    if (cur != INF) {  // document `cur` is complete: filter
                        bool pass = true;
                        if (tv) {
                            if (F < Flo) {
                                pass = false;
                            } else if (cnt == 1) {
                                bool tie = lj == cj && lw == cw;
                                if (!tie && F <= Fhi) {
                                    double tfd = (double)(lw >> 8);
                                    double Sx = __ddiv_rn(__dmul_rn(tfd, h->s0d[lj]), __dadd_rn(tfd, p.s1d[lw & 0xFFu]));
                                    if (Sx == Sk) {
                                        cj = lj;
                                        cw = lw;
                                        tie = true;
                                    } else if (Sx < Sk) {
                                        pass = false;
                                    }
                                }
                                if (tie && cur > dk) pass = false;  // equal score, larger doc id: cannot enter
                            }
                        }
                        if (pass) {
                            int idx = atomicAdd((int *)&ctrl->qn, 1);
                            if (idx < C::QC) {
                                q_doc[idx] = cur;
                            } else {
                                ctrl->stall = 1;  // queue full: retry this document after the drain
                                break;
                            }
                        }
                    }
                    cur = dmin;
                    F = 0.f;
                    cnt = 0;
                }
                if (dmin == INF) {  // every sub-run of my bucket is exhausted (also the empty-bucket case)
                    done = true;
                    break;
                }
                // consume the head posting of run jm: Cache::evaluate (bm25.rs:355-358) in f32
                {
                    float tff = (float)(wm >> 8);
                    F += __fdividef(tff * s0m, tff + s1f[wm & 0xFFu]);
                    cnt++;
                    lj = (uint32_t)jm;
                    lw = wm;
                }
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
            // ---- phase B: exact f64 score of the survivors, reference operation order ----
            const int nqueue = min(ctrl->qn, C::QC);
            const int stalled = ctrl->stall;
            for (int c = tid; c < nqueue; c += T) {
                const uint32_t doc = q_doc[c];
                if (p.allow && !((p.allow[doc >> 3] >> (doc & 7u)) & 1u)) continue;  // filter(payload), search.rs:230
                double Sx = 0.0;
                for (uint32_t j = 0; j < m; ++j) {
                    const uint32_t a = h->run_off[j];
                    uint32_t l = 0, r = h->run_len[j];
                    while (l < r) {
                        uint32_t mid = (l + r) >> 1;
                        if (st[a + mid].doc < doc) l = mid + 1;
                        else r = mid;
                    }
                    if (l < h->run_len[j]) {
                        Posting v = st[a + l];
                        if (v.doc == doc) {
                            double tfd = (double)(v.w >> 8);
                            Sx = __dadd_rn(Sx, __ddiv_rn(__dmul_rn(tfd, h->s0d[j]), __dadd_rn(tfd, p.s1d[v.w & 0xFFu])));
                        }
                    }
                }
                if (!tv || Sx > Sk || (Sx == Sk && doc < dk)) {
                    int idx = atomicAdd((int *)&ctrl->pool_n, 1);
                    pool_s[idx] = (uint64_t)__double_as_longlong(Sx);
                    pool_d[idx] = doc;
                }
            }
            cbar<T>();
            // ---- pool upkeep ----
            const int pn = ctrl->pool_n;
            const bool fin = !stalled && (flags & FLAG_LAST);
            const int slack = (int)k > 64 ? (int)k : 64;
            const bool need = pn > 0 && (fin || pn > C::PC - C::QC || pn >= (int)k + slack);
            if (tid == 0) {
                ctrl->qn = 0;
                ctrl->stall = 0;
            }
            if (need) {
                int n2 = 2;
                while (n2 < pn) n2 <<= 1;
                pool_sort<T>(pool_s, pool_d, pn, n2, tid);
                if (tid == 0) {
                    int nn = pn < (int)k ? pn : (int)k;
                    ctrl->pool_n = nn;
                    if (nn == (int)k) {
                        double sk = __longlong_as_double((long long)pool_s[k - 1]);
                        ctrl->Sk = sk;
                        ctrl->dk = pool_d[k - 1];
                        ctrl->Flo = __double2float_rd(sk * (1.0 - kEps));
                        ctrl->Fhi = __double2float_ru(sk * (1.0 + kEps));
                        ctrl->thr_valid = 1;
                    }
                }
            }
            cbar<T>();
            if (!stalled) break;
        }
        if (flags & FLAG_LAST) {  // Results::into_sorted_vec (search.rs:281): the pool is sorted, best first
            const int n = ctrl->pool_n;
            const size_t base = (size_t)qid * k;
            for (int i = tid; i < (int)k; i += T) {
                uint32_t d = INF;
                double sc = 0.0;
                if (i < n) {
                    d = pool_d[i];
                    sc = __longlong_as_double((long long)pool_s[i]);
                }
                p.out_doc[base + i] = d;
                p.out_score[base + i] = (float)sc;
                if (p.out_score64) p.out_score64[base + i] = sc;
                if (p.out_payload) {
                    uint16_t a = 0, b = 0, c = 0;
                    if (i < n) {
                        a = p.payload[(size_t)d * 3 + 0];
                        b = p.payload[(size_t)d * 3 + 1];
                        c = p.payload[(size_t)d * 3 + 2];
                    }
                    p.out_payload[(base + i) * 3 + 0] = a;
                    p.out_payload[(base + i) * 3 + 1] = b;
                    p.out_payload[(base + i) * 3 + 2] = c;
                }
            }
            if (tid == 0) p.out_n[qid] = (uint32_t)n;
            cbar<T>();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
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
        uint64_t *full = (uint64_t *)(smem + S::off_bar);
        uint64_t *empty = full + C::STAGES;
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], C::T / 32);
        }
        mbar_fence_init();
    }
    for (int i = tid; i < 256; i += C::THREADS) ((float *)(smem + S::off_s1f))[i] = p.s1f[i];
    __syncthreads();
    if (tid >= C::T) producer<C>(p, smem, tid - C::T);
    else consumer<C>(p, smem, tid);
}

// ---------------------------------------------------------------------------------------------
// bm25::evaluate (crates/bm25/src/evaluate.rs:22-74): one thread per (document, query) pair.
__global__ void k_evaluate(uint32_t n_pairs, const uint32_t *__restrict__ d_off, const uint32_t *__restrict__ d_terms,
                           const uint32_t *__restrict__ d_tfs, const uint32_t *__restrict__ q_off,
                           const uint32_t *__restrict__ q_terms, const uint32_t *__restrict__ fn_len,
                           const double *__restrict__ idf, const double *__restrict__ s1d, const uint32_t *__restrict__ df,
                           uint32_t n_terms, double k1, double *__restrict__ out) {
    uint32_t pidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (pidx >= n_pairs) return;
    uint32_t a = d_off[pidx], b = d_off[pidx + 1];
    uint64_t len = 0;  // Document::length(): saturating Σ tf (vector.rs:77-83)
    for (uint32_t i = a; i < b; ++i) {
        len += d_tfs[i];
        if (len > 0xFFFFFFFFull) len = 0xFFFFFFFFull;
    }
    int lo = 0, hi = 256;  // length_to_fieldnorm, bm25.rs:278-283
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (fn_len[mid] <= (uint32_t)len) lo = mid + 1;
        else hi = mid;
    }
    const int fn = lo - 1;
    uint32_t cursor = a;
    double result = 0.0;
    const double k1p1 = __dadd_rn(k1, 1.0);
    for (uint32_t qi = q_off[pidx]; qi < q_off[pidx + 1]; ++qi) {
        uint32_t key = q_terms[qi];
        while (cursor < b && d_terms[cursor] < key) cursor++;
        if (!(cursor < b && d_terms[cursor] == key)) continue;
        if (key >= n_terms || df[key] == 0) continue;  // address_tokens::read → None
        double tfd = (double)d_tfs[cursor];
        double tfv = __ddiv_rn(__dmul_rn(tfd, k1p1), __dadd_rn(tfd, s1d[fn]));  // bm25.rs:291-295
        result = __dadd_rn(result, __dmul_rn(idf[key], tfv));
    }
    out[pidx] = result;
}

}  // namespace

// =============================================================================================
// Host side
// =============================================================================================

static const int kClasses[] = {1, 2, 3, 4, 8, 16, 32};
static const int kNumClasses = 7;

struct Group {
    int M = 0;
    uint32_t nq = 0;
    std::vector<uint32_t> h_ids, h_off, h_terms;
    uint32_t *d_ids = nullptr, *d_off = nullptr, *d_terms = nullptr;
    int *d_counter = nullptr;
};

struct bm25x_batch {
    bm25x_index *ix = nullptr;
    uint32_t nq = 0, k = 0;
    Group groups[kNumClasses];
    uint8_t *d_allow = nullptr;
    uint32_t *d_out_doc = nullptr;
    float *d_out_score = nullptr;
    double *d_out_score64 = nullptr;
    uint16_t *d_out_payload = nullptr;
    uint32_t *d_out_n = nullptr;
    uint64_t postings = 0, qterms = 0;
    uint32_t live = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<void *> allocs;
    void *last_stream = nullptr;
};

template <class C>
static int launch_class(const bm25x_index *ix, SearchParams &sp, cudaStream_t stream) {
    using S = Smem<C>;
    static bool configured[64] = {false};
    auto kern = k_search<C>;
    if (!configured[ix->device & 63]) {
        BM25X_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::total));
        configured[ix->device & 63] = true;
    }
    int per_sm = 0;
    BM25X_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, C::THREADS, S::total));
    if (per_sm < 1) {
        bm25x_set_error("k_search<M=%d> does not fit on an SM (smem %zu)", C::M, (size_t)S::total);
        return BM25X_ERR_CUDA;
    }
    uint32_t grid = (uint32_t)(per_sm * ix->sm_count);
    if (grid > sp.nq) grid = sp.nq;
    kern<<<grid, C::THREADS, S::total, stream>>>(sp);
    BM25X_CUDA_TRY(cudaGetLastError());
    return BM25X_OK;
}

template <typename T>
static int batch_alloc(bm25x_batch *b, T **p, size_t n) {
    BM25X_CUDA_TRY(cudaMalloc((void **)p, sizeof(T) * (n ? n : 1)));
    b->allocs.push_back((void *)*p);
    return BM25X_OK;
}

extern "C" void bm25x_batch_destroy(bm25x_batch *b) {
    if (!b) return;
    cudaSetDevice(b->ix->device);
    for (void *p : b->allocs) cudaFree(p);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    delete b;
}

#define BTRY(x)                      \
    do {                             \
        int _rc = (x);               \
        if (_rc != BM25X_OK) {       \
            bm25x_batch_destroy(b);  \
            return _rc;              \
        }                            \
    } while (0)

extern "C" int bm25x_batch_prepare(bm25x_index *ix, uint32_t nq, const uint32_t *q_off, const uint32_t *q_terms,
                                   uint32_t k, const uint8_t *allow, bm25x_batch **out) {
    if (!ix || !out || (nq && (!q_off || (!q_terms && q_off[nq] != 0)))) {
        bm25x_set_error("bm25x_batch_prepare: null argument");
        return BM25X_ERR_INVALID;
    }
    *out = nullptr;
    if (k == 0) {
        bm25x_set_error("number of needed rows is set to 0");  // scanners/default.rs:114-116
        return BM25X_ERR_LIMIT_ZERO;
    }
    if (k > BM25X_MAX_K) {
        bm25x_set_error("bm25x_batch_prepare: k=%u > BM25X_MAX_K=%d", k, BM25X_MAX_K);
        return BM25X_ERR_UNSUPPORTED;
    }
    bm25x_batch *b = new bm25x_batch();
    b->ix = ix;
    b->nq = nq;
    b->k = k;
    for (int c = 0; c < kNumClasses; ++c) {
        b->groups[c].M = kClasses[c];
        b->groups[c].h_off.push_back(0);
    }
    // canonicalise: sort + dedup (datatype/tsvector.rs:96-105), drop unknown tokens (search.rs:55-62)
    std::vector<uint32_t> tmp;
    const uint32_t T = ix->d.n_terms;
    for (uint32_t i = 0; i < nq; ++i) {
        if (q_off[i + 1] < q_off[i]) {
            bm25x_set_error("bm25x_batch_prepare: q_off not monotone at %u", i);
            delete b;
            return BM25X_ERR_INVALID;
        }
        tmp.assign(q_terms + q_off[i], q_terms + q_off[i + 1]);
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        size_t m = 0;
        for (uint32_t t : tmp)
            if (t < T && ix->h_df[t] != 0) tmp[m++] = t;
        tmp.resize(m);
        if (m == 0) continue;
        if (m > BM25X_MAX_QUERY_TERMS) {
            bm25x_set_error("bm25x_batch_prepare: query %u has %zu live terms > %d", i, m, BM25X_MAX_QUERY_TERMS);
            delete b;
            return BM25X_ERR_UNSUPPORTED;
        }
        int c = 0;
        while (kClasses[c] < (int)m) ++c;
        Group &g = b->groups[c];
        g.h_ids.push_back(i);
        for (uint32_t t : tmp) {
            g.h_terms.push_back(t);
            b->postings += ix->h_df[t];
        }
        g.h_off.push_back((uint32_t)g.h_terms.size());
        g.nq++;
        b->qterms += m;
        b->live++;
    }
    cudaError_t e = cudaSetDevice(ix->device);
    if (e != cudaSuccess) {
        bm25x_set_error("cudaSetDevice: %s", cudaGetErrorString(e));
        delete b;
        return BM25X_ERR_CUDA;
    }
    cudaStream_t st = ix->stream;
    for (int c = 0; c < kNumClasses; ++c) {
        Group &g = b->groups[c];
        if (!g.nq) continue;
        BTRY(batch_alloc(b, &g.d_ids, g.h_ids.size()));
        BTRY(batch_alloc(b, &g.d_off, g.h_off.size()));
        BTRY(batch_alloc(b, &g.d_terms, g.h_terms.size()));
        BTRY(batch_alloc(b, &g.d_counter, 1));
        cudaMemcpyAsync(g.d_ids, g.h_ids.data(), 4 * g.h_ids.size(), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(g.d_off, g.h_off.data(), 4 * g.h_off.size(), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(g.d_terms, g.h_terms.data(), 4 * g.h_terms.size(), cudaMemcpyHostToDevice, st);
    }
    if (allow) {
        size_t nb = ((size_t)ix->d.n_docs + 7) / 8;
        BTRY(batch_alloc(b, &b->d_allow, nb));
        cudaMemcpyAsync(b->d_allow, allow, nb, cudaMemcpyHostToDevice, st);
    }
    size_t slots = (size_t)nq * k;
    if (slots == 0) slots = 1;
    BTRY(batch_alloc(b, &b->d_out_doc, slots));
    BTRY(batch_alloc(b, &b->d_out_score, slots));
    BTRY(batch_alloc(b, &b->d_out_score64, slots));
    BTRY(batch_alloc(b, &b->d_out_payload, slots * 3));
    BTRY(batch_alloc(b, &b->d_out_n, nq));
    e = cudaMemsetAsync(b->d_out_n, 0, 4 * (size_t)(nq ? nq : 1), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_out_doc, 0xFF, 4 * (slots ? slots : 1), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_out_score, 0, 4 * (slots ? slots : 1), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_out_score64, 0, 8 * (slots ? slots : 1), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_out_payload, 0, 6 * (slots ? slots : 1), st);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev1);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        bm25x_set_error("bm25x_batch_prepare: %s", cudaGetErrorString(e));
        bm25x_batch_destroy(b);
        return BM25X_ERR_CUDA;
    }
    *out = b;
    return BM25X_OK;
}

extern "C" int bm25x_batch_run(bm25x_batch *b, void *stream_v, bm25x_search_stats *stats) {
    if (!b) {
        bm25x_set_error("bm25x_batch_run: null batch");
        return BM25X_ERR_INVALID;
    }
    bm25x_index *ix = b->ix;
    BM25X_CUDA_TRY(cudaSetDevice(ix->device));
    cudaStream_t st = stream_v ? (cudaStream_t)stream_v : ix->stream;
    b->last_stream = (void *)st;
    const DeviceIndex &d = ix->d;
    uint32_t launches = 0;
    if (stats) BM25X_CUDA_TRY(cudaEventRecord(b->ev0, st));
    for (int c = 0; c < kNumClasses; ++c) {
        Group &g = b->groups[c];
        if (!g.nq) continue;
        SearchParams sp;
        sp.post = d.post;
        sp.post_off = d.post_off;
        sp.df = d.df;
        sp.blk_off = d.blk_off;
        sp.blk = d.blk;
        sp.s0f = d.s0f;
        sp.s0d = d.s0d;
        sp.s1d = d.s1d;
        sp.s1f = d.s1f;
        sp.payload = d.payload;
        sp.n_docs = d.n_docs;
        sp.q_ids = g.d_ids;
        sp.q_off = g.d_off;
        sp.q_terms = g.d_terms;
        sp.nq = g.nq;
        sp.k = b->k;
        sp.allow = b->d_allow;
        sp.work_counter = g.d_counter;
        sp.out_doc = b->d_out_doc;
        sp.out_score = b->d_out_score;
        sp.out_score64 = b->d_out_score64;
        sp.out_payload = b->d_out_payload;
        sp.out_n = b->d_out_n;
        BM25X_CUDA_TRY(cudaMemsetAsync(g.d_counter, 0, sizeof(int), st));
        int rc = BM25X_OK;
        switch (g.M) {
            case 1: rc = launch_class<KCfg<1>>(ix, sp, st); break;
            case 2: rc = launch_class<KCfg<2>>(ix, sp, st); break;
            case 3: rc = launch_class<KCfg<3>>(ix, sp, st); break;
            case 4: rc = launch_class<KCfg<4>>(ix, sp, st); break;
            case 8: rc = launch_class<KCfg<8>>(ix, sp, st); break;
            case 16: rc = launch_class<KCfg<16>>(ix, sp, st); break;
            default: rc = launch_class<KCfg<32>>(ix, sp, st); break;
        }
        if (rc != BM25X_OK) return rc;
        launches++;
    }
    if (stats) {
        BM25X_CUDA_TRY(cudaEventRecord(b->ev1, st));
        BM25X_CUDA_TRY(cudaEventSynchronize(b->ev1));
        float ms = 0.f;
        BM25X_CUDA_TRY(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
        memset(stats, 0, sizeof(*stats));
        stats->kernel_ms = ms;
        stats->postings = b->postings;
        stats->bytes_algo = 8ull * b->postings + 8ull * (uint64_t)b->live * b->k + 16ull * b->qterms;
        stats->launches = launches;
        stats->queries = b->live;
    }
    return BM25X_OK;
}

extern "C" int bm25x_batch_fetch(bm25x_batch *b, uint32_t *out_doc, float *out_score, double *out_score64,
                                 uint16_t *out_payload, uint32_t *out_n) {
    if (!b) {
        bm25x_set_error("bm25x_batch_fetch: null batch");
        return BM25X_ERR_INVALID;
    }
    BM25X_CUDA_TRY(cudaSetDevice(b->ix->device));
    cudaStream_t st = b->last_stream ? (cudaStream_t)b->last_stream : b->ix->stream;
    size_t slots = (size_t)b->nq * b->k;
    if (out_doc) BM25X_CUDA_TRY(cudaMemcpyAsync(out_doc, b->d_out_doc, 4 * slots, cudaMemcpyDeviceToHost, st));
    if (out_score) BM25X_CUDA_TRY(cudaMemcpyAsync(out_score, b->d_out_score, 4 * slots, cudaMemcpyDeviceToHost, st));
    if (out_score64) BM25X_CUDA_TRY(cudaMemcpyAsync(out_score64, b->d_out_score64, 8 * slots, cudaMemcpyDeviceToHost, st));
    if (out_payload) BM25X_CUDA_TRY(cudaMemcpyAsync(out_payload, b->d_out_payload, 6 * slots, cudaMemcpyDeviceToHost, st));
    if (out_n) BM25X_CUDA_TRY(cudaMemcpyAsync(out_n, b->d_out_n, 4 * (size_t)b->nq, cudaMemcpyDeviceToHost, st));
    BM25X_CUDA_TRY(cudaStreamSynchronize(st));
    return BM25X_OK;
}

extern "C" int bm25x_search_batch(bm25x_index *ix, uint32_t nq, const uint32_t *q_off, const uint32_t *q_terms,
                                  uint32_t k, const uint8_t *allow, uint32_t *out_doc, float *out_score,
                                  double *out_score64, uint16_t *out_payload, uint32_t *out_n,
                                  bm25x_search_stats *stats) {
    bm25x_batch *b = nullptr;
    int rc = bm25x_batch_prepare(ix, nq, q_off, q_terms, k, allow, &b);
    if (rc != BM25X_OK) return rc;
    bm25x_search_stats local;
    rc = bm25x_batch_run(b, nullptr, stats ? stats : &local);
    if (rc == BM25X_OK) rc = bm25x_batch_fetch(b, out_doc, out_score, out_score64, out_payload, out_n);
    bm25x_batch_destroy(b);
    return rc;
}

// ---------------------------------------------------------------------------------------------
uint32_t bm25x_fieldnorm_to_length(uint8_t fn);

extern "C" int bm25x_evaluate_batch(bm25x_index *ix, uint32_t n_pairs, const uint32_t *d_off, const uint32_t *d_terms,
                                    const uint32_t *d_tfs, const uint32_t *q_off, const uint32_t *q_terms, double *out) {
    if (!ix || (n_pairs && (!d_off || !q_off || !out))) {
        bm25x_set_error("bm25x_evaluate_batch: null argument");
        return BM25X_ERR_INVALID;
    }
    if (n_pairs == 0) return BM25X_OK;
    BM25X_CUDA_TRY(cudaSetDevice(ix->device));
    const uint32_t nd = d_off[n_pairs], nqt = q_off[n_pairs];
    const uint32_t T = ix->d.n_terms;
    // idf table (bm25.rs:285-289) with the host libm, like the reference's f64::ln
    std::vector<double> h_idf(T ? T : 1);
    for (uint32_t t = 0; t < T; ++t)
        h_idf[t] = log(((double)ix->d.n_docs + 1.0) / ((double)ix->h_df[t] + 0.5));
    uint32_t h_fn[256];
    for (int f = 0; f < 256; ++f) h_fn[f] = bm25x_fieldnorm_to_length((uint8_t)f);
    uint32_t *g_doff = nullptr, *g_dt = nullptr, *g_df = nullptr, *g_qoff = nullptr, *g_qt = nullptr, *g_fn = nullptr;
    double *g_idf = nullptr, *g_out = nullptr;
    int rc = BM25X_OK;
    cudaError_t e = cudaSuccess;
    auto A = [&](void **p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes ? bytes : 4);
    };
    A((void **)&g_doff, 4 * ((size_t)n_pairs + 1));
    A((void **)&g_dt, 4 * (size_t)nd);
    A((void **)&g_df, 4 * (size_t)nd);
    A((void **)&g_qoff, 4 * ((size_t)n_pairs + 1));
    A((void **)&g_qt, 4 * (size_t)nqt);
    A((void **)&g_fn, 4 * 256);
    A((void **)&g_idf, 8 * (size_t)(T ? T : 1));
    A((void **)&g_out, 8 * (size_t)n_pairs);
    auto H = [&](void *dst, const void *src, size_t bytes) {
        if (e == cudaSuccess && bytes) e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    };
    H(g_doff, d_off, 4 * ((size_t)n_pairs + 1));
    H(g_dt, d_terms, 4 * (size_t)nd);
    H(g_df, d_tfs, 4 * (size_t)nd);
    H(g_qoff, q_off, 4 * ((size_t)n_pairs + 1));
    H(g_qt, q_terms, 4 * (size_t)nqt);
    H(g_fn, h_fn, sizeof(h_fn));
    H(g_idf, h_idf.data(), 8 * (size_t)T);
    if (e == cudaSuccess) {
        k_evaluate<<<(n_pairs + 127) / 128, 128>>>(n_pairs, g_doff, g_dt, g_df, g_qoff, g_qt, g_fn, g_idf, ix->d.s1d,
                                                   ix->d.df, T, ix->k1, g_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, g_out, 8 * (size_t)n_pairs, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        bm25x_set_error("bm25x_evaluate_batch: %s", cudaGetErrorString(e));
        rc = BM25X_ERR_CUDA;
    }
    cudaFree(g_doff);
    cudaFree(g_dt);
    cudaFree(g_df);
    cudaFree(g_qoff);
    cudaFree(g_qt);
    cudaFree(g_fn);
    cudaFree(g_idf);
    cudaFree(g_out);
    return rc;
}
