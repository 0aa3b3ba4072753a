#!/bin/bash
# final validation of the round: full GPU suite, smoke, default bench (both arms), ncu captures of the two dominant kernels, launch list
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $OUT/r2z_tests.log 2>&1
tail -5 $OUT/r2z_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2z_smoke.log 2>&1; tail -2 $OUT/r2z_smoke.log
T0=$(date +%s); python bench.py > $OUT/r2z_bench.json 2> $OUT/r2z_bench.err; echo "bench.py wall $(( $(date +%s) - T0 )) s"
T0=$(date +%s); python bench.py --impl reference > $OUT/r2z_bench_reference.json 2> $OUT/r2z_bench_reference.err; echo "bench.py --impl reference wall $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
for n in ("r2z_bench", "r2z_bench_reference"):
    try:
        j = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(j["value"], 1), "e2e", j.get("e2e", {}).get("value"), "ms/step", round(j["ms_per_step"], 2), "cpu", (j.get("cpu_baseline") or {}).get("value"), "parity", j.get("parity_sample"), "recall", j.get("recall"))
    except Exception as e:
        print(n, "unreadable", e)
PY
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --recall-queries 0 --no-other-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_walk_kernel -s 2 -c 1 -o $OUT/r2z_prof_walk -f python bench.py $ARGS > /dev/null 2> $OUT/r2z_ncu_walk.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kw_search_kernel -s 1 -c 1 -o $OUT/r2z_prof_kw -f python bench.py $ARGS > /dev/null 2> $OUT/r2z_ncu_kw.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $OUT/r2z_launches.csv python bench.py $ARGS > $OUT/r2z_bench_under_ncu.json 2> $OUT/r2z_ncu_launches.log
ls -la $OUT | grep r2z_
