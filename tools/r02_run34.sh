#!/bin/bash
# device ART walks: larger chunks of 2-typo searches (fewer, fuller frontier launches) — end-to-end effect
OUT=gpurun_out
mkdir -p $OUT
for v in "64 16777216" "256 33554432" "512 67108864"; do
  set -- $v
  export TSGPU_ART_CHUNK2=$1 TSGPU_ART_ITEMS=$2
  timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-other-configs --recall-queries 0 --no-cpu-baseline > $OUT/r2m_bench_$1.json 2> $OUT/r2m_bench_$1.err; echo "chunk2=$1 rc=$?"
  python - <<PY
import json
j = json.loads(open("gpurun_out/r2m_bench_$1.json").read().strip().splitlines()[-1])
print("chunk2=$1 e2e", round(j["e2e"]["value"]), "ms", round(j["e2e"]["ms_per_step"], 1), "small", round(j["latency_ms"]["small"]["p50"], 1), {k: round(v, 1) for k, v in j["host_rounds_per_step"].items() if k.startswith("ms_")})
PY
done
