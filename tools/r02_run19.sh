#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
run() { # name batch env...
  name=$1; batch=$2; shift; shift
  env "$@" python bench.py --batch $batch --steps 3 --warmup 2 --no-cpu-baseline --recall-queries 0 --no-other-configs > $OUT/r2s_$name.json 2> $OUT/r2s_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    j = json.loads(open(f"gpurun_out/r2s_{n}.json").read().strip().splitlines()[-1])
    w = j["work_per_step"]
    print(n, "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ms/step", round(j["ms_per_step"], 2), "knn", round(j["device_ms_isolated"]["knn"], 2), "kw", round(j["device_ms_isolated"]["kw_search"], 2),
          "probes/dist", round(w["knn_table_probes"] / max(w["knn_dist"], 1), 2), "walks", j.get("knn_walks", {}).get("expanded_max"), "small p50", round(j["latency_ms"]["small"]["p50"], 2))
except Exception as e:
    print(n, "unreadable", e)
PY
}
run ut4096 4096 TSGPU_KW_UNIT_TARGET=4096
run ut16384 4096 TSGPU_KW_UNIT_TARGET=16384
run ut32768 4096 TSGPU_KW_UNIT_TARGET=32768
ARGS="--steps 2 --warmup 1 --no-cpu-baseline --recall-queries 0 --no-other-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_walk_kernel -s 2 -c 1 -o $OUT/r2s_prof_walk -f python bench.py $ARGS > /dev/null 2> $OUT/r2s_ncu_walk.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kw_search_kernel -s 1 -c 1 -o $OUT/r2s_prof_kw -f python bench.py $ARGS > /dev/null 2> $OUT/r2s_ncu_kw.log
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $OUT/r2s_launches.csv python bench.py $ARGS > $OUT/r2s_bench_under_ncu.json 2> $OUT/r2s_ncu_launches.log
ls -la $OUT | grep r2s_
