#!/bin/bash
# parity of the end-to-end path with and without the host layer's staging pool (debugging aid TSHOST_NO_STAGING)
OUT=gpurun_out
mkdir -p $OUT
for v in pool nostaging; do
  if [ $v = nostaging ]; then export TSHOST_NO_STAGING=1; else unset TSHOST_NO_STAGING; fi
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-other-configs --recall-queries 0 > $OUT/r2j_bench_$v.json 2> $OUT/r2j_bench_$v.err; echo "$v rc=$?"
  python - <<PY
import json
j = json.loads(open("gpurun_out/r2j_bench_$v.json").read().strip().splitlines()[-1])
print("$v", j.get("parity_sample"), j.get("parity_resolved_vs_cpu", {}).get("identical_topk"), "e2e", j["e2e"]["value"])
PY
done
