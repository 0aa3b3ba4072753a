#!/bin/bash
# staging pool + ART output reuse: host-layer tests on the GPU, then the driver's bench line
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_zz_art_gpu.py tests/test_cpp_host.py tests/test_host_batched.py -x -q -m gpu -p no:cacheprovider > $OUT/r2l_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/r2l_tests.log
T0=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2l_bench.json 2> $OUT/r2l_bench.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[-1])
    print("value", round(j["value"], 1), "e2e", j["e2e"]["value"], "e2e ms", j["e2e"]["ms_per_step"], "cpu", (j.get("cpu_baseline") or {}).get("value"),
          "parity", (j.get("parity_sample") or {}).get("identical_topk"), "lat", j["latency_ms"], "rounds", j.get("host_rounds_per_step"))
    oc = j.get("other_configs") or {}
    print("kw typo", oc.get("keyword10m_typo"))
except Exception as e:
    print("bench unreadable", e)
PY
tail -3 $OUT/r2l_bench.err
