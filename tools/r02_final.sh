#!/bin/bash
# final validation of round 2: full GPU suite, smoke, the driver's bench line (both arms), ncu captures of the dominant kernels + the tensor-core scan, launch list
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s); timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $OUT/r2z_tests.log 2>&1; echo "gpu tests rc=$? wall $(( $(date +%s) - T0 )) s"; tail -3 $OUT/r2z_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r2z_smoke.log
T0=$(date +%s); timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2z_bench.json 2> $OUT/r2z_bench.err; echo "bench rc=$? wall $(( $(date +%s) - T0 )) s"
T0=$(date +%s); timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/r2z_bench_reference.json 2> $OUT/r2z_bench_reference.err; echo "reference arm rc=$? wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
for n in ("r2z_bench", "r2z_bench_reference"):
    try:
        j = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(j["value"], 1), "e2e", round(j.get("e2e", {}).get("value", 0), 1), "ms/step", round(j["ms_per_step"], 2), "cpu", (j.get("cpu_baseline") or {}).get("value"),
              "parity", j.get("parity_sample"), "recall", j.get("knn_recall_at_100"), "roof", (j.get("roofline") or {}).get("kernel"), (j.get("roofline") or {}).get("frac"),
              [(r["kernel"], round(r["frac"], 3)) for r in j.get("roofline_other", [])], "flat_tc", ((j.get("other_configs") or {}).get("flat_scan_tensor") or {}).get("roofline"))
    except Exception as e:
        print(n, "unreadable", e)
PY
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --recall-queries 0 --no-other-configs"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_walk_kernel -s 2 -c 1 -o $OUT/r2z_prof_walk -f python bench.py $ARGS > /dev/null 2> $OUT/r2z_ncu_walk.log; echo "ncu walk rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kw_search_kernel -s 1 -c 1 -o $OUT/r2z_prof_kw -f python bench.py $ARGS > /dev/null 2> $OUT/r2z_ncu_kw.log; echo "ncu kw rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $OUT/r2z_launches.csv python bench.py $ARGS > $OUT/r2z_bench_under_ncu.json 2> $OUT/r2z_ncu_launches.log; echo "ncu launches rc=$?"
timeout 200 ./tests/cpp/flat_tc_check 200000 > $OUT/flat_tc_check.log 2>&1; echo "flat_tc check rc=$?"; grep "ids=200000\|PASSED\|FAILED" $OUT/flat_tc_check.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_tc_kernel -s 12 -c 1 -o $OUT/r2z_prof_flat_tc -f ./tests/cpp/flat_tc_check 200000 > $OUT/r2z_ncu_flat_tc.log 2>&1; echo "ncu flat_tc rc=$?"
ls -la $OUT | grep "r2z_\|flat_tc"
