#!/bin/bash
# r3d: final validation of the round (seeded kernel with the holder loads in flight together, sliced bm25x_search_batch),
# then the secondary workloads at size
mkdir -p gpurun_out
BM25X_SEED_FORCE=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zx_stress.py -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r3d_gate.log
cat gpurun_out/r3d_gate.log
grep -q "failed\|error" gpurun_out/r3d_gate.log && { echo "GATE FAILED: stopping"; exit 1; }
bash tools/final_validate.sh r3d
O=gpurun_out/r3d
timeout 600 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu-baseline > ${O}_bench_c5_n1.json 2> ${O}_bench_c5_n1.err; echo "c5 rc=$?"
python - <<PY
import json
for w in ("c2", "c4_pruned_100k", "c5_n1"):
    try:
        l = json.loads(open("${O}_bench_%s.json" % w).read().strip().splitlines()[-1])
        print(w, round(l["value"]), "q/s", round(l["ms_per_step"], 3), "ms frac", round(l["roofline"]["frac"], 4), "e2e",
              round(l["e2e"]["value"]), "skipped", round(l["roofline"]["skipped_frac"], 3), l["config"].get("gen_s"), l["config"].get("index_build_s"))
    except Exception as e:
        print(w, "failed", e)
PY
