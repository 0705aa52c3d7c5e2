#!/bin/bash
# Round-2 GPU pass g: secondary workloads at size (C4 pruned 100k / exhaustive 4k, C2, C5 on one GPU).
mkdir -p gpurun_out
O=gpurun_out/r2p
timeout 400 python bench.py --workload c4 --queries 100000 --steps 3 --warmup 1 > ${O}_bench_c4_pruned_100k.json 2> ${O}_bench_c4_pruned_100k.err; echo "c4 pruned rc=$?"
timeout 400 python bench.py --workload c4 --queries 4000 --no-prune --no-cpu-baseline --steps 2 --warmup 1 > ${O}_bench_c4_exhaustive_4k.json 2> ${O}_bench_c4_exhaustive_4k.err; echo "c4 exhaustive rc=$?"
timeout 200 python bench.py --workload c2 --steps 10 --warmup 3 > ${O}_bench_c2.json 2> ${O}_bench_c2.err; echo "c2 rc=$?"
timeout 900 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu-baseline > ${O}_bench_c5_n1.json 2> ${O}_bench_c5_n1.err; echo "c5 rc=$?"
python - <<PY
import json
for w in ("c4_pruned_100k", "c4_exhaustive_4k", "c2", "c5_n1"):
    try:
        l = json.loads(open("${O}_bench_%s.json" % w).read().strip().splitlines()[-1])
        print(w, round(l["value"]), "q/s", round(l["ms_per_step"], 3), "ms frac", round(l["roofline"]["frac"], 4), "e2e",
              round(l["e2e"]["value"]), "cpu", l["cpu_baseline"] and round(l["cpu_baseline"]["value"]),
              "skipped", round(l["roofline"]["skipped_frac"], 3), l["config"].get("gen_s"), l["config"].get("index_build_s"))
    except Exception as e:
        print(w, "failed", e)
PY
tail -3 ${O}_bench_c5_n1.err
