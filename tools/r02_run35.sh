#!/bin/bash
# end-to-end leg: requests in flight (8 = default) vs 12 / 16
OUT=gpurun_out
mkdir -p $OUT
for d in 4 6; do
  timeout 600 python bench.py --gpus 1 --steps 24 --warmup 5 --no-other-configs --recall-queries 0 --no-cpu-baseline --e2e-depth $d > $OUT/r2o_bench_d$d.json 2> $OUT/r2o_bench_d$d.err; echo "depth=$d rc=$?"
  python - <<PY
import json
j = json.loads(open("gpurun_out/r2o_bench_d$d.json").read().strip().splitlines()[-1])
print("depth=$d e2e", round(j["e2e"]["value"]), "ms", round(j["e2e"]["ms_per_step"], 1), "batch p50", round(j["latency_ms"]["batch"]["p50"]), {k: round(v, 1) for k, v in j["host_rounds_per_step"].items() if k.startswith("ms_")})
PY
done
