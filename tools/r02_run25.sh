#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
run() { # name env...
  name=$1; shift
  env "$@" python bench.py --steps 8 --warmup 4 --no-cpu-baseline --recall-queries 0 --no-other-configs --e2e-depth 8 --e2e-threads 32 > $OUT/r2x_$name.json 2> $OUT/r2x_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/r2x_{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), round(j["e2e"]["ms_per_step"], 1), "small p50", round(j["latency_ms"]["small"]["p50"], 2), "p99", round(j["latency_ms"]["small"]["p99"], 2),
          {k: round(v, 1) for k, v in j["host_rounds_per_step"].items() if k.startswith("ms_")})
except Exception as e:
    print(n, "unreadable", e)
PY
}
run c128 TSGPU_ART_CHUNK2=128
run c256 TSGPU_ART_CHUNK2=256 TSGPU_ART_ITEMS=33554432
run c512 TSGPU_ART_CHUNK2=512 TSGPU_ART_ITEMS=67108864
run c32 TSGPU_ART_CHUNK2=32
tail -3 $OUT/r2x_c128.err
