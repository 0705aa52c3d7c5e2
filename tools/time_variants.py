#!/usr/bin/env python
"""Times tuning variants of libbm25x.so (tools/build_variants.sh) on one GPU and checks that they all return the same
bits.  Each variant runs in its own process (BM25X_LIBRARY); the workload is bench.py's C3 corpus with (a) the C3
queries, top-10, (b) the same, top-100, (c) 100k queries of 1..8 terms (the C5 mix).  Prints one line per variant and
writes gpurun_out/variants.json.  Not part of the product or the tests."""
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(docs):
    import numpy as np
    import _pkg
    m = _pkg.load()
    t0 = time.time()
    c = m.synth_corpus(0xB25C0DE0 + 3, docs, 100_000, 128, 128, 0.0)
    ix = m.Index.from_corpus(c)
    out = {"build_s": round(time.time() - t0, 1)}
    for name, (nq, tmin, tmax, k) in {"c3": (100_000, 3, 3, 10), "c3k100": (100_000, 3, 3, 100),
                                      "c5mix": (100_000, 1, 8, 10)}.items():
        q_off, q_terms = m.synth_queries(0xB25C0DE0 + 1003, nq, 100_000, tmin, tmax, c.post_off, 0.0)
        b = ix.prepare(q_off, q_terms, k)
        for _ in range(3):
            b.run()
        ms = sorted(b.run().kernel_ms for _ in range(7))
        r = b.fetch()
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(r["doc"]).tobytes())
        h.update(np.ascontiguousarray(r["score64"]).tobytes())
        out[name] = {"ms_median": round(ms[len(ms) // 2], 3), "ms_min": round(ms[0], 3), "sha": h.hexdigest()[:16]}
        b.close()
    print("VARIANT " + json.dumps(out), flush=True)


def main():
    if sys.argv[1] == "--one":
        return one(int(sys.argv[2]))
    docs = int(os.environ.get("VAR_DOCS", 10_000_000))
    res = {}
    for name in sys.argv[1:]:
        lib = os.path.join(ROOT, "vectorchord-bm25_b200", "variants", f"libbm25x_{name}.so")
        env = dict(os.environ, BM25X_LIBRARY=lib)
        try:
            p = subprocess.run([sys.executable, __file__, "--one", str(docs)], env=env, capture_output=True, text=True,
                               timeout=75)
            line = [l for l in p.stdout.splitlines() if l.startswith("VARIANT ")]
            res[name] = json.loads(line[-1][8:]) if line else {"error": (p.stderr or p.stdout)[-400:]}
        except subprocess.TimeoutExpired:
            res[name] = {"error": "timeout"}
        print(name, json.dumps(res[name]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "variants.json"), "w"), indent=1)
    ok = [n for n in res if "c3" in res[n]]
    if ok:
        ref = res[ok[0]]
        for n in ok:
            same = all(res[n][w]["sha"] == ref[w]["sha"] for w in ("c3", "c3k100", "c5mix"))
            print(f"{n:12s} c3 {res[n]['c3']['ms_median']:8.3f} ms  k100 {res[n]['c3k100']['ms_median']:8.3f} ms  "
                  f"c5mix {res[n]['c5mix']['ms_median']:8.3f} ms  {'same bits as ' + ok[0] if same else 'RESULTS DIFFER'}")


if __name__ == "__main__":
    main()
