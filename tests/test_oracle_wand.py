"""Differential test inside the oracle: the Block-max WAND restatement (search.rs:28-282) must
return the same top-k as the exhaustive f64 scorer modulo score ties — the reference's own fuzz
test makes the same comparison (index scan vs seq scan, tests/fuzz:217-303) with a looser bar."""
import numpy as np
import pytest

from util_parity import check_topk


@pytest.mark.parametrize("cfg", [
    dict(seed=11, n=1000, vocab=1000, lmin=32, lmax=32, zipf=0.0),      # BASELINE config 1 shape
    dict(seed=12, n=3000, vocab=300, lmin=1, lmax=200, zipf=0.0),       # varied lengths → fieldnorms vary
    dict(seed=13, n=5000, vocab=2000, lmin=16, lmax=64, zipf=1.0),      # Zipf: dense head lists, tf > 1
    dict(seed=14, n=700, vocab=50, lmin=5, lmax=400, zipf=1.1),         # many blocks per term, heavy ties
])
def test_wand_equals_exhaustive(orc, cfg):
    c = orc.Corpus.synth(cfg["seed"], cfg["n"], cfg["vocab"], cfg["lmin"], cfg["lmax"], cfg["zipf"])
    ix = orc.OracleIndex(c)
    q_off, q_terms = orc.gen_queries(cfg["seed"] + 1000, 60, cfg["vocab"], 1, 8, ix.df, cfg["zipf"])
    for k in (1, 10, 100):
        for i in range(len(q_off) - 1):
            q = q_terms[q_off[i]:q_off[i + 1]]
            od, os_, _ = ix.search_exhaustive(q, k)
            st = orc.WandStats()
            wd, ws = ix.search_wand(q, k, stats=st)

            def score_of(d):
                full_d, full_s, _ = ix.search_exhaustive(q, cfg["n"])
                return float(full_s[list(full_d).index(d)])
            check_topk(wd, ws, od, os_, rtol=1e-12, score_of=score_of, what=f"cfg{cfg['seed']} q{i} k{k}")


def test_wand_prunes(orc):
    c = orc.Corpus.synth(21, 20000, 500, 20, 60, 1.0)
    ix = orc.OracleIndex(c)
    q_off, q_terms = orc.gen_queries(22, 20, 500, 4, 4, ix.df, 1.0)
    total = touched = 0
    for i in range(20):
        q = q_terms[q_off[i]:q_off[i + 1]]
        st = orc.WandStats()
        ix.search_wand(q, 10, stats=st)
        total += sum(ix.df(int(t)) for t in q)
        touched += st.postings_touched
    assert touched < total  # block-max skipping is doing something


def test_filter_bitmap(orc):
    c = orc.Corpus.synth(31, 2000, 100, 10, 30, 0.0)
    ix = orc.OracleIndex(c)
    allow = np.zeros((2000 + 7) // 8, dtype=np.uint8)
    rng = np.random.default_rng(3)
    keep = rng.random(2000) < 0.3
    for d in np.nonzero(keep)[0]:
        allow[d >> 3] |= 1 << (d & 7)
    q = [3, 17, 42]
    od, os_, _ = ix.search_exhaustive(q, 20, allow=allow)
    wd, ws = ix.search_wand(q, 20, allow=allow)
    assert all(keep[d] for d in od)
    check_topk(wd, ws, od, os_, rtol=1e-12)


def test_unknown_and_duplicate_terms(orc):
    c = orc.Corpus.synth(41, 500, 64, 8, 8, 0.0)
    ix = orc.OracleIndex(c)
    a = ix.search_exhaustive([5, 9], 10)
    b = ix.search_exhaustive([9, 5, 5, 1000000, 9], 10)   # dedup + unknown dropped (search.rs:55-62)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert len(ix.search_exhaustive([1000000], 10)[0]) == 0
    assert len(ix.search_wand([1000000], 10)[0]) == 0
