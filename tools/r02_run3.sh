#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --recall-queries 0 > $OUT/r2c_bench.json 2> $OUT/r2c_bench.err
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2c_bench.json").read().strip().splitlines()[-1])
print("value", round(j["value"]), "ms/step", round(j["ms_per_step"], 2), "iso", j.get("device_ms_isolated"))
print(j.get("knn_walks"))
PY
tail -3 $OUT/r2c_bench.err
