#!/usr/bin/env python
"""Times library variants / kernel generations side by side on ONE GPU and checks that all return the same bits.

  python tools/time_variants.py NAME[@opt=v,...] ...   NAME = variants/libbm25x_NAME.so ("main" = the in-tree library),
                                                       opt=v = bm25x_index_set_option of that run (e.g. main@seed=0,twophase=1)
Env: VAR_DOCS (default 10M), VAR_WORKLOADS (comma list of c3,c3k100,c5mix,c2,c4).  Every variant runs in its own process.
"""
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WL = {  # name: (queries, tmin, tmax, k, zipf)
    "c3": (100_000, 3, 3, 10, 0.0), "c3k100": (100_000, 3, 3, 100, 0.0), "c5mix": (100_000, 1, 8, 10, 0.0),
    "c2": (100_000, 1, 1, 10, 0.0), "c3k1000": (20_000, 3, 3, 1000, 0.0),
    # VAR_CORPUS=zipf (the C4 corpus: Zipf(1) term frequencies): 8-term queries, pruning on / off (exhaustive)
    "c4": (4_000, 8, 8, 10, 1.0), "c4np": (400, 8, 8, 10, 1.0), "c4mix": (20_000, 1, 8, 10, 1.0),
}


def one(docs, workloads):
    import numpy as np
    import _pkg
    m = _pkg.load()
    t0 = time.time()
    zipf_corpus = os.environ.get("VAR_CORPUS", "uniform") == "zipf"
    c = m.synth_corpus(0xB25C0DE0 + (4 if zipf_corpus else 3), docs, 100_000, 128, 128, 1.0 if zipf_corpus else 0.0)
    ix = m.Index.from_corpus(c)
    for kv in filter(None, os.environ.get("VAR_OPTS", "").split(",")):  # NAME@opt=v,opt=v: index options of this variant
        name, _, v = kv.partition("=")
        ix.set_option(name, int(v))
    out = {"build_s": round(time.time() - t0, 1)}
    for name in workloads:
        nq, tmin, tmax, k, zipf = WL[name]
        assert (zipf > 0) == zipf_corpus, "workload and VAR_CORPUS do not match"
        q_off, q_terms = m.synth_queries(0xB25C0DE0 + 1003 + (1 if zipf_corpus else 0), nq, 100_000, tmin, tmax,
                                         c.post_off, zipf)
        ix.set_option("prune", 0 if name.endswith("np") else 1)
        b = ix.prepare(q_off, q_terms, k)
        for _ in range(3):
            b.run()
        runs = [b.run() for _ in range(7)]
        ms = sorted(r.kernel_ms for r in runs)
        r = b.fetch()
        h = hashlib.sha256()
        h.update(np.ascontiguousarray(r["doc"]).tobytes())
        h.update(np.ascontiguousarray(r["score64"]).tobytes())
        out[name] = {"ms_median": round(ms[len(ms) // 2], 3), "ms_min": round(ms[0], 3), "sha": h.hexdigest()[:16],
                     "fetched_ratio": round(runs[0].postings_fetched / max(1, runs[0].postings), 3)}
        b.close()
    print("VARIANT " + json.dumps(out), flush=True)


def main():
    workloads = os.environ.get("VAR_WORKLOADS", "c3,c3k100,c5mix").split(",")
    if sys.argv[1] == "--one":
        return one(int(sys.argv[2]), workloads)
    docs = int(os.environ.get("VAR_DOCS", 10_000_000))
    res = {}
    for spec in sys.argv[1:]:
        name, _, gen = spec.partition("@")
        env = dict(os.environ)
        env["VAR_OPTS"] = gen
        if name != "main":
            env["BM25X_LIBRARY"] = os.path.join(ROOT, "vectorchord-bm25_b200", "variants", f"libbm25x_{name}.so")
        try:
            p = subprocess.run([sys.executable, __file__, "--one", str(docs)], env=env, capture_output=True, text=True,
                               timeout=int(os.environ.get("VAR_TIMEOUT", 120)))
            line = [l for l in p.stdout.splitlines() if l.startswith("VARIANT ")]
            res[spec] = json.loads(line[-1][8:]) if line else {"error": (p.stderr or p.stdout)[-600:]}
        except subprocess.TimeoutExpired:
            res[spec] = {"error": "timeout"}
        print(spec, json.dumps(res[spec]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = os.environ.get("VAR_TAG", "variants")
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"{tag}.json"), "w"), indent=1)
    ok = [n for n in res if workloads[0] in res[n]]
    if ok:
        ref = res[ok[0]]
        for n in ok:
            same = all(res[n][w]["sha"] == ref[w]["sha"] for w in workloads)
            print(f"{n:16s} " + "  ".join(f"{w} {res[n][w]['ms_median']:8.3f} ms" for w in workloads) +
                  f"  {'same bits as ' + ok[0] if same else 'RESULTS DIFFER'}")


if __name__ == "__main__":
    main()
