"""Batching broker (include/bm25x_broker.h) on the CPU: the ring / worker / scatter logic against a stub
backend with the signature of bm25x_search_batch.  (The same broker over a real index handle: tests/test_gpu_parity.py.)"""
import threading

import numpy as np
import pytest

import _pkg


@pytest.fixture(scope="module")
def bm():
    mod = _pkg.load()
    mod.build_library()
    mod.load_library()
    return __import__(mod.__name__ + ".bm25x", fromlist=["x"])


def _rows(terms, k):
    """What the stub engine answers for one query: a deterministic, totally ordered list (best first)."""
    n = min(k, 3 + 2 * len(terms))
    base = int(np.sum(terms, dtype=np.uint64) % 100000)
    docs = np.array([base + 7 * r for r in range(n)], np.uint32)
    scores = np.array([1000.0 - r - 0.001 * len(terms) for r in range(n)], np.float64)
    return docs, scores


class Stub:
    def __init__(self, fail_when_k=None):
        self.calls = []   # (nq, k) per backend call
        self.fail_when_k = fail_when_k

    def __call__(self, ctx, nq, q_off, q_terms, k, out_doc, out_score, out_score64, out_payload, out_n):
        self.calls.append((nq, k))
        if self.fail_when_k is not None and k == self.fail_when_k:
            return 2
        for i in range(nq):
            terms = np.array([q_terms[j] for j in range(q_off[i], q_off[i + 1])], np.uint32)
            docs, scores = _rows(terms, k)
            for r in range(len(docs)):
                out_doc[i * k + r] = int(docs[r])
                out_score[i * k + r] = float(scores[r])
                out_score64[i * k + r] = float(scores[r])
                if out_payload:
                    for c in range(3):
                        out_payload[(i * k + r) * 3 + c] = (int(docs[r]) + c) & 0xFFFF
            out_n[i] = len(docs)
        return 0


def test_single_caller_roundtrip(bm):
    stub = Stub()
    br = bm.Broker(backend=stub, max_wait_us=50)
    docs, scores, pay = br.search([5, 9, 11], 4, want_payload=True)
    wd, ws = _rows(np.array([5, 9, 11], np.uint32), 4)
    assert np.array_equal(docs, wd) and np.array_equal(scores, ws)
    assert np.array_equal(pay[:, 1], (wd + 1) & 0xFFFF)
    d0, s0 = br.search([], 10)          # no token: the engine's answer for an empty query
    assert len(d0) == 3
    st = br.stats()
    assert st.requests == 2 and st.batches == 2 and st.rejected == 0
    br.close()


def test_concurrent_callers_are_coalesced_and_answered_exactly(bm):
    stub = Stub()
    br = bm.Broker(backend=stub, max_batch=64, max_wait_us=20000, ring_slots=16)  # small ring: callers also queue up for space
    rng = np.random.default_rng(11)
    work = [(rng.integers(0, 5000, size=int(rng.integers(0, 9)), dtype=np.uint32), int(rng.choice([1, 3, 10, 32, 33, 100, 128, 200, 1000])))
            for _ in range(240)]
    got = [None] * len(work)
    errs = []

    def client(lo, hi):
        try:
            for i in range(lo, hi):
                got[i] = br.search(work[i][0], work[i][1])
        except Exception as e:  # pragma: no cover
            errs.append(e)

    threads = [threading.Thread(target=client, args=(i * 10, i * 10 + 10)) for i in range(24)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for (terms, k), (docs, scores) in zip(work, got):
        wd, ws = _rows(terms, k)   # served from a batch run at a larger limit: the first k rows are the top-k
        assert np.array_equal(docs, wd) and np.array_equal(scores, ws)
    st = br.stats()
    assert st.requests == len(work)
    assert st.batches < len(work) and st.max_batch_seen > 1, "no coalescing happened"
    assert st.max_batch_seen <= 64
    # one backend call per limit class and batch: its limit is the largest one of the class's requests
    assert all(k in (1, 3, 10, 32, 33, 100, 128, 200, 1000) for _, k in stub.calls)
    assert sum(nq for nq, _ in stub.calls) == len(work)
    br.close()


def test_rejections_do_not_reach_the_batch(bm):
    stub = Stub()
    br = bm.Broker(backend=stub, max_wait_us=50)
    with pytest.raises(bm.Bm25xError) as e:
        br.search([1, 2], 0)
    assert e.value.code == 5 and "number of needed rows is set to 0" in str(e.value)
    with pytest.raises(bm.Bm25xError) as e:
        br.search([1], bm.MAX_K + 1)
    assert e.value.code == 4
    with pytest.raises(bm.Bm25xError) as e:
        br.search(list(range(bm.MAX_QUERY_TERMS + 1)), 10)
    assert e.value.code == 4
    assert stub.calls == [] and br.stats().rejected == 3
    docs, _ = br.search([1, 2], 5)       # the broker still works
    assert len(docs) == 5
    br.close()


def test_backend_failure_reaches_every_caller_of_that_batch_only(bm):
    stub = Stub(fail_when_k=100)
    br = bm.Broker(backend=stub, max_batch=32, max_wait_us=30000)
    out = {}

    def client(name, k):
        try:
            out[name] = br.search([3, 4], k)
        except bm.Bm25xError as e:
            out[name] = e

    ts = [threading.Thread(target=client, args=(f"bad{i}", 100)) for i in range(3)] + \
         [threading.Thread(target=client, args=(f"ok{i}", 10)) for i in range(3)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for i in range(3):
        assert isinstance(out[f"bad{i}"], bm.Bm25xError) and out[f"bad{i}"].code == 2
        assert not isinstance(out[f"ok{i}"], Exception) and len(out[f"ok{i}"][0]) == 7
    br.close()


def test_destroy_answers_what_is_queued(bm):
    stub = Stub()
    br = bm.Broker(backend=stub, max_batch=1000, max_wait_us=2_000_000)  # the worker would wait 2 s for more requests
    res = []
    t = threading.Thread(target=lambda: res.append(br.search([8], 2)))
    t.start()
    import time
    time.sleep(0.2)
    br.close()             # stop: the queued request is answered first
    t.join(timeout=5)
    assert not t.is_alive() and len(res) == 1 and len(res[0][0]) == 2
