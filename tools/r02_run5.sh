#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_hnsw_build_gpu.py -q -m gpu -x -p no:cacheprovider -k "knn or vector or hybrid or build" > $OUT/r2e_tests.log 2>&1
tail -15 $OUT/r2e_tests.log
timeout 900 python tools/bench_build_gpu.py 1000000 768 8192 200 > $OUT/r2e_build_1m.json 2> $OUT/r2e_build_1m.err
tail -2 $OUT/r2e_build_1m.json; tail -3 $OUT/r2e_build_1m.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --recall-queries 0 > $OUT/r2e_bench.json 2> $OUT/r2e_bench.err
TSGPU_KNN_HEAP=1 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --recall-queries 0 > $OUT/r2e_bench_heap.json 2> $OUT/r2e_bench_heap.err
python - <<'PY'
import json
for n in ("r2e_bench", "r2e_bench_heap"):
    try:
        j = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 2), "iso", j.get("device_ms_isolated"), j.get("knn_walks"))
    except Exception as e:
        print(n, "unreadable", e)
PY
tail -3 $OUT/r2e_bench.err
