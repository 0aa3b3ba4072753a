#!/bin/bash
# round 2, GPU call 1: A/B default vs TSGPU_REG_SCORE=1, ncu full of the regscore kw kernel, ART walk timings (both forms)
OUT=gpurun_out
mkdir -p $OUT
ARGS="--steps 6 --warmup 3 --no-cpu-baseline --recall-queries 0"
python bench.py $ARGS > $OUT/r2a_default.json 2> $OUT/r2a_default.err
TSGPU_REG_SCORE=1 python bench.py $ARGS > $OUT/r2a_regscore.json 2> $OUT/r2a_regscore.err
python - <<'PY'
import json
for n in ("default", "regscore"):
    try:
        j = json.loads(open(f"gpurun_out/r2a_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 2), "iso", j.get("device_ms_isolated"), "small", j["latency_ms"]["small"])
    except Exception as e:
        print(n, "unreadable:", e)
PY
TSGPU_REG_SCORE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:kw_search_kernel -s 1 -c 1 -o $OUT/r2a_prof_kw_regscore -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --recall-queries 0 > /dev/null 2> $OUT/r2a_ncu.log
timeout 300 python tools/bench_art_gpu.py > $OUT/r2a_art_dfs.json 2> $OUT/r2a_art_dfs.err
TSGPU_ART_MODE=frontier timeout 300 python tools/bench_art_gpu.py > $OUT/r2a_art_frontier.json 2> $OUT/r2a_art_frontier.err
tail -1 $OUT/r2a_art_dfs.json; tail -1 $OUT/r2a_art_frontier.json
ls -la $OUT | grep r2a
