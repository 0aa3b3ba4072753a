#!/bin/bash
# state of the build on the B200: tensor-core flat scan v2 (check + ncu), the whole GPU suite, the driver's bench line
OUT=gpurun_out
mkdir -p $OUT
timeout 200 ./tests/cpp/flat_tc_check 200000 > $OUT/flat_tc_check.log 2>&1; echo "check rc=$?"; grep -c "^ok" $OUT/flat_tc_check.log; grep "ids=200000\|FAIL\|PASSED\|FAILED" $OUT/flat_tc_check.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:flat_tc_kernel -s 12 -c 1 -o $OUT/r2fu_prof_flat_tc -f ./tests/cpp/flat_tc_check 200000 > $OUT/r2fu_ncu.log 2>&1; echo "ncu rc=$?"
T0=$(date +%s); timeout 2400 python -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=8 > $OUT/r2fu_tests.log 2>&1; echo "pytest rc=$? wall $(( $(date +%s) - T0 )) s"; tail -14 $OUT/r2fu_tests.log
T0=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2fu_bench.json 2> $OUT/r2fu_bench.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2fu_bench.json").read().strip().splitlines()[-1])
    print("value", round(j["value"], 1), "e2e", j["e2e"]["value"], "ms/step", round(j["ms_per_step"], 2), "e2e ms", j["e2e"]["ms_per_step"], "cpu", (j.get("cpu_baseline") or {}).get("value"),
          "parity", j.get("parity_sample"), "recall", j.get("knn_recall_at_100"), "lat", j.get("latency_ms"), "rounds", j.get("host_rounds_per_step"))
    print("roof", j["roofline"]["kernel"], j["roofline"]["frac"], [(r["kernel"], r["frac"]) for r in j.get("roofline_other", [])])
    print("other", json.dumps(j.get("other_configs"))[:1500])
except Exception as e:
    print("bench unreadable", e)
PY
tail -5 $OUT/r2fu_bench.err
