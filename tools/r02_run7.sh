#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_hnsw_build_gpu.py -q -m gpu -x -p no:cacheprovider -k "knn or vector or hybrid or batched" > $OUT/r2g_tests.log 2>&1
tail -12 $OUT/r2g_tests.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --recall-queries 0 > $OUT/r2g_bench.json 2> $OUT/r2g_bench.err
python - <<'PY'
import json
for n in ("r2g_bench",):
    try:
        j = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 2), "iso", j.get("device_ms_isolated"), j.get("knn_walks"))
        print([(r["kernel"], round(r["frac"], 3), round(r["ms"], 2)) for r in [j["roofline"]] + j["roofline_other"]], j["latency_ms"]["small"])
    except Exception as e:
        print(n, "unreadable", e)
PY
tail -4 $OUT/r2g_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_walk_kernel -s 2 -c 1 -o $OUT/r2g_prof_walk -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --recall-queries 0 > /dev/null 2> $OUT/r2g_ncu.log
