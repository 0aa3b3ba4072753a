#!/bin/bash
# round 2, GPU call 2: kNN kernel with the hashed visited set — parity tests, bench, launch list, ncu full of the walk
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "knn or vector or hybrid" > $OUT/r2b_tests.log 2>&1
tail -5 $OUT/r2b_tests.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --recall-queries 0 > $OUT/r2b_bench.json 2> $OUT/r2b_bench.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2b_bench.json").read().strip().splitlines()[-1])
print("value", round(j["value"]), "ms/step", round(j["ms_per_step"], 2), "iso", j.get("device_ms_isolated"), "small", j["latency_ms"]["small"])
print([ (r["kernel"], round(r["frac"],3), r["ms"]) for r in [j["roofline"]] + j["roofline_other"]])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 2 -c 1 -o $OUT/r2b_prof_knn -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --recall-queries 0 > /dev/null 2> $OUT/r2b_ncu.log
ls -la $OUT | grep r2b
