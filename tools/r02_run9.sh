#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider > $OUT/r2i_tests.log 2>&1
tail -8 $OUT/r2i_tests.log
timeout 1500 python bench.py --steps 4 --warmup 3 --recall-queries 64 > $OUT/r2i_bench.json 2> $OUT/r2i_bench.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2i_bench.json").read().strip().splitlines()[-1])
    print("value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "e2e_resolved", round(j["e2e_resolved"]["value"]), "ms/step", round(j["ms_per_step"], 2), "e2e ms", round(j["e2e"]["ms_per_step"], 2))
    print("iso", j.get("device_ms_isolated"))
    print("host rounds", j.get("host_rounds_per_step"))
    print("cpu", j.get("cpu_baseline"))
    print("parity", j.get("parity_sample"), j.get("parity_resolved_vs_cpu"))
    print("lat", j["latency_ms"], "recall", j.get("knn_recall_at_100"))
    print("e2e xfer", j["e2e"]["h2d_bytes_per_step"], j["e2e"]["d2h_bytes_per_step"], j["e2e"]["device_calls_per_step"])
except Exception as e:
    print("unreadable", e)
PY
tail -12 $OUT/r2i_bench.err
