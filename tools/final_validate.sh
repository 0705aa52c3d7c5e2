#!/bin/bash
# Round-end validation on one B200 (run through gpurun): GPU tests, smoke, the default bench line, the ncu captures
# that profiles/ summarises, and the secondary workloads.  Outputs: gpurun_out/${TAG}_*.
TAG=${1:-r1b}
mkdir -p gpurun_out
O=gpurun_out/$TAG
timeout 200 python -m pytest tests -q -m gpu 2>&1 | tail -5 > ${O}_pytest.log; cat ${O}_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -1 ${O}_smoke.log
timeout 150 python bench.py > ${O}_bench_c3.json 2> ${O}_bench_c3.err; echo "bench rc=$?"
timeout 170 ncu --set full --clock-control none --import-source on -k regex:k_search_wq -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
    python bench.py --steps 1 --warmup 2 --no-cpu-baseline --queries 20000 > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 170 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > ${O}_ncu_b.log 2>&1; echo "ncu list rc=$?"
timeout 120 python bench.py --workload c4 --cpu-seconds 8 > ${O}_bench_c4.json 2> ${O}_bench_c4.err; echo "c4 rc=$?"
timeout 120 python bench.py --workload c5 --docs 10000000 --cpu-seconds 6 > ${O}_bench_c5shape.json 2> ${O}_bench_c5shape.err; echo "c5 rc=$?"
timeout 60 python bench.py --workload c2 --cpu-seconds 5 > ${O}_bench_c2.json 2> ${O}_bench_c2.err; echo "c2 rc=$?"
python - <<PY
import json
for w in ("c3", "c4", "c5shape", "c2"):
    try:
        l = json.loads(open("${O}_bench_%s.json" % w).read().strip().splitlines()[-1])
        print(w, round(l["value"]), "q/s", round(l["ms_per_step"], 3), "ms frac", round(l["roofline"]["frac"], 4), "e2e",
              round(l["e2e"]["value"]), "cpu", l["cpu_baseline"] and round(l["cpu_baseline"]["value"]), "top100", l.get("top100"),
              "skipped", round(l["roofline"]["skipped_frac"], 3), l["clocks"])
    except Exception as e:
        print(w, "failed", e)
PY
