#!/bin/bash
# usage: bash tools/r02_scale.sh N   (inside gpurun --gpus N)
N=$1
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 \
    --no-cpu-baseline --no-other-configs --recall-queries 0 > $OUT/r2zs_bench_${N}gpu.json 2> $OUT/r2zs_bench_${N}gpu.err
tail -3 $OUT/r2zs_bench_${N}gpu.err
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
j = json.loads(open(f"gpurun_out/r2zs_bench_{n}gpu.json").read().strip().splitlines()[-1])
print("N", n, "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 2), "e2e", round(j["e2e"]["value"]), round(j["e2e"]["ms_per_step"], 1), "collective", j.get("collective"), "knn", j["device_ms_isolated"])
PY
