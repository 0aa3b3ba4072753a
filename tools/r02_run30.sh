#!/bin/bash
# device ART walks with warp-aggregated frontier allocation: parity tests, stand-alone timing (with the phase split), the driver's bench line
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_zz_art_gpu.py -x -q -m gpu -p no:cacheprovider > $OUT/r2g_art_tests.log 2>&1; echo "art tests rc=$?"; tail -3 $OUT/r2g_art_tests.log
TSGPU_ART_TIMING=1 timeout 600 python tools/bench_art_gpu.py > $OUT/r2g_art_gpu.json 2> $OUT/r2g_art_gpu.err; echo "art bench rc=$?"; cat $OUT/r2g_art_gpu.json; grep "tsgpu art" $OUT/r2g_art_gpu.err | tail -8
T0=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2g_bench.json 2> $OUT/r2g_bench.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
    print("value", round(j["value"], 1), "e2e", j["e2e"]["value"], "e2e ms", j["e2e"]["ms_per_step"], "cpu", (j.get("cpu_baseline") or {}).get("value"),
          "parity", (j.get("parity_sample") or {}).get("identical_topk"), "lat small", j["latency_ms"]["small"], "rounds", j.get("host_rounds_per_step"))
    oc = j.get("other_configs") or {}
    print("kw typo", oc.get("keyword10m_typo")); print("flat tc", oc.get("flat_scan_tensor"))
except Exception as e:
    print("bench unreadable", e)
PY
tail -3 $OUT/r2g_bench.err
