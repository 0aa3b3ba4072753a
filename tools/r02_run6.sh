#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "contains" > $OUT/r2f_tests.log 2>&1
tail -3 $OUT/r2f_tests.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --recall-queries 64 > $OUT/r2f_bench.json 2> $OUT/r2f_bench.err
TSGPU_KNN_HEAP=1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --recall-queries 0 > $OUT/r2f_bench_heap.json 2> $OUT/r2f_bench_heap.err
python - <<'PY'
import json
for n in ("r2f_bench", "r2f_bench_heap"):
    try:
        j = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 2), "iso", j.get("device_ms_isolated"), j.get("knn_walks"), "recall", j.get("knn_recall_at_100"))
        print([(r["kernel"], round(r["frac"], 3), round(r["ms"], 2)) for r in [j["roofline"]] + j["roofline_other"]])
    except Exception as e:
        print(n, "unreadable", e)
PY
cat $OUT/r2f_bench.err | tail -8
