#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_zz_art_gpu.py tests/test_host_batched.py tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider -k "art or batched or hybrid" > $OUT/r2k_tests.log 2>&1
tail -6 $OUT/r2k_tests.log
timeout 1500 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --recall-queries 0 > $OUT/r2k_bench.json 2> $OUT/r2k_bench.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1])
    print("value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "e2e_resolved", round(j["e2e_resolved"]["value"]), "ms/step", round(j["ms_per_step"], 2), "e2e ms", round(j["e2e"]["ms_per_step"], 2))
    print("host rounds", j.get("host_rounds_per_step"))
    print("lat", j["latency_ms"])
except Exception as e:
    print("unreadable", e)
PY
tail -5 $OUT/r2k_bench.err
