#!/usr/bin/env python
"""CPU model of the chunk planner of k_search_wq (bm25x_search_wq.cuh: plan_prefetch / plan_chunk, exact-quota loads)
for one query class, to size things before spending GPU time: how many postings a chunk loads and consumes, and how
long the possible-duplicate list of phase B gets, under different quota rules and tag-map configurations.

It reproduces: quota split (floor + leftover to the first terms = shipped; largest remainder = BM25X_BALANCED), one-block
quotas for the first two chunks, the quarter-width first window, window end = smallest doc id of the first posting
not loaded, no re-load of a term whose next posting lies past the window end, tag maps with the kernel's two hashes.
It does NOT model time.  Cross-check against hardware: `postings_streamed / postings_exhaustive` of the bench line
(C3, shipped rule) is 1.665; the model gives the `loaded/consumed` column.

  python tools/sim_planner.py [--queries 400] [--terms 3] [--docs 10000000] [--df 12800]
"""
import argparse

import numpy as np

H1, H2 = 0x9E3779B1, 0x85EBCA77


def quotas(dfs, cb, balanced):
    m = len(dfs)
    extra = cb - m
    tot = int(sum(dfs))
    share = [(extra * int(d)) // tot for d in dfs]
    left = extra - sum(share)
    if balanced:
        rem = [(extra * int(d)) % tot for d in dfs]
        order = sorted(range(m), key=lambda j: (-rem[j], j))
        bonus = set(order[:left])
        return [1 + share[j] + (1 if j in bonus else 0) for j in range(m)]
    return [1 + share[j] + (1 if j < left else 0) for j in range(m)]


def victims(runs, log_slots, two):
    """Postings whose tag lost its slot(s) to a later run (phase B's possible duplicates)."""
    if two:
        s1 = [((r.astype(np.uint64) * H1) & 0xFFFFFFFF) >> (32 - (log_slots - 1)) for r in runs]
        s2 = [((r.astype(np.uint64) * H2) & 0xFFFFFFFF) >> (32 - (log_slots - 1)) for r in runs]
    else:
        s1 = [((r.astype(np.uint64) * H1) & 0xFFFFFFFF) >> (32 - log_slots) for r in runs]
    nd = 0
    for j in range(len(runs) - 1):
        later1 = np.concatenate(s1[j + 1:]) if j + 1 < len(runs) else np.zeros(0, np.uint64)
        lost1 = np.isin(s1[j], later1)
        if two:
            later2 = np.concatenate(s2[j + 1:])
            nd += int(np.count_nonzero(lost1 & np.isin(s2[j], later2)))
        else:
            nd += int(np.count_nonzero(lost1))
    return nd


def simulate(lists, cb, balanced, configs):
    m = len(lists)
    dfs = [len(x) for x in lists]
    qf = quotas(dfs, cb, balanced)
    gpos = [0] * m
    next_doc = [0] * m
    lo, chunk = 0, 0
    stats = {"chunks": 0, "loaded": 0, "consumed": 0, "nd": {c: [] for c in configs}}
    INF = 1 << 40
    while True:
        act = [gpos[j] < dfs[j] for j in range(m)]
        if not any(act):
            break
        q = [(1 if chunk < 2 else qf[j]) if act[j] else 0 for j in range(m)]
        gs = [gpos[j] & ~1 for j in range(m)]
        endp = [min(gs[j] + q[j] * 128, dfs[j]) for j in range(m)]
        prop = [int(lists[j][gs[j] + q[j] * 128]) if act[j] and gs[j] + q[j] * 128 < dfs[j] else INF for j in range(m)]
        hi = min(prop)
        if chunk == 0 and hi != INF:
            hi = lo + max(1, (hi - lo) >> 2)
        runs = []
        for j in range(m):
            if not act[j] or (hi != INF and next_doc[j] >= hi):
                continue
            seg = lists[j][gpos[j]:endp[j]]
            stats["loaded"] += ((endp[j] - gs[j] + 1) & ~1)
            e = int(np.searchsorted(seg, hi, side="left")) if hi != INF else len(seg)
            runs.append(seg[:e])
            stats["consumed"] += e
            gpos[j] += e
            next_doc[j] = int(lists[j][gpos[j]]) if gpos[j] < endp[j] else 0
        for c in configs:
            stats["nd"][c].append(victims(runs, *configs[c]) if len(runs) > 1 else 0)
        stats["chunks"] += 1
        chunk += 1
        lo = hi
        if hi == INF:
            break
    return stats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=400)
    ap.add_argument("--terms", type=int, default=3)
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--df", type=int, default=12_800, help="mean document frequency of a query term (C3: 12 800)")
    ap.add_argument("--cbmul", type=int, default=2)
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    configs = {"1 map 4 KiB": (12, False), "2 maps 4 KiB": (12, True), "1 map 8 KiB": (13, False),
               "2 maps 8 KiB": (13, True)}
    print(f"{a.terms}-term queries, df ~ {a.df}, {a.docs} docs, {a.cbmul} blocks per term and chunk, {a.queries} queries")
    for balanced in (False, True):
        tot = {"chunks": 0, "loaded": 0, "consumed": 0, "nd": {c: [] for c in configs}}
        for _ in range(a.queries):
            dfs = rng.binomial(a.docs, a.df / a.docs, a.terms)
            lists = [np.sort(rng.choice(a.docs, int(d), replace=False)).astype(np.uint32) for d in dfs]
            s = simulate(lists, a.cbmul * a.terms, balanced, configs)
            for k in ("chunks", "loaded", "consumed"):
                tot[k] += s[k]
            for c in configs:
                tot["nd"][c] += s["nd"][c]
        name = "largest-remainder quotas (BM25X_BALANCED=1)" if balanced else "floor + leftover-to-first quotas (shipped)"
        print(f"\n{name}\n  chunks/query {tot['chunks'] / a.queries:7.1f}   consumed/chunk {tot['consumed'] / tot['chunks']:7.1f}"
              f"   loaded/consumed {tot['loaded'] / tot['consumed']:.3f}")
        for c in configs:
            nd = np.array(tot["nd"][c])
            print(f"  {c:14s} possible duplicates per chunk: mean {nd.mean():6.1f}  p99 {np.percentile(nd, 99):5.0f}"
                  f"  P(>32) {np.mean(nd > 32):6.3f}  P(>64) {np.mean(nd > 64):6.4f}")


if __name__ == "__main__":
    main()
