#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_zz_art_gpu.py tests/test_host_batched.py -q -m gpu -x -p no:cacheprovider > $OUT/r2w_tests.log 2>&1
tail -4 $OUT/r2w_tests.log
timeout 300 python tools/bench_art_gpu.py > $OUT/r2w_art_frontier.json 2> $OUT/r2w_art_frontier.err
cat $OUT/r2w_art_frontier.json | cut -c1-600
run() { # name depth threads
  python bench.py --steps 8 --warmup 4 --no-cpu-baseline --recall-queries 0 --no-other-configs --e2e-depth $2 --e2e-threads $3 > $OUT/r2w_$1.json 2> $OUT/r2w_$1.err
  python - "$1" <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/r2w_{n}.json").read().strip().splitlines()[-1])
    print(n, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), round(j["e2e"]["ms_per_step"], 1), "small p50", round(j["latency_ms"]["small"]["p50"], 2), "p99", round(j["latency_ms"]["small"]["p99"], 2),
          {k: round(v, 1) for k, v in j["host_rounds_per_step"].items()})
except Exception as e:
    print(n, "unreadable", e)
PY
}
run d4t64 4 64
run d6t32 6 32
run d8t32 8 32
run d8t16 8 16
tail -3 $OUT/r2w_d4t64.err
