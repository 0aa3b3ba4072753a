#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --recall-queries 0 > $OUT/r2l_bench2.json 2> $OUT/r2l_bench2.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2l_bench2.json").read().strip().splitlines()[-1])
    print("N=2 value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 2), "e2e ms", round(j["e2e"]["ms_per_step"], 2), j.get("collective"), j["scaling"])
except Exception as e:
    print("unreadable", e)
PY
tail -8 $OUT/r2l_bench2.err
