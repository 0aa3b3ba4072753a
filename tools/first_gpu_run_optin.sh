#!/bin/bash
# The first GPU minutes of the next round, in one gpurun call (≈ 12 min of box time):
#   tools/gpurun_retry.sh 1500 'bash tools/first_gpu_run_optin.sh'
# 1. the two opt-in kernels' parity tests for real (--runxfail turns their xfail(strict=False) into pass / fail);
# 2. A/B of the keyword kernel: default vs TSGPU_REG_SCORE=1 (register-resident scoring + single-field build), same seed;
# 3. one ncu --set full capture of the opt-in kernel for the local-memory / issue-slot comparison with profiles/r01c_ncu_full.md.
# Everything lands in gpurun_out/optin_*.
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_zz_regscore_gpu.py tests/test_zz_art_gpu.py -q -m gpu --runxfail -p no:cacheprovider > $OUT/optin_tests.log 2>&1
tail -3 $OUT/optin_tests.log
ARGS="--steps 6 --warmup 3 --no-cpu-baseline"
python bench.py $ARGS > $OUT/optin_bench_default.json 2> $OUT/optin_bench_default.err
TSGPU_REG_SCORE=1 python bench.py $ARGS > $OUT/optin_bench_regscore.json 2> $OUT/optin_bench_regscore.err
python - <<'PY'
import json
for n in ("default", "regscore"):
    try:
        j = json.loads(open(f"gpurun_out/optin_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(j["value"]), "ms/step", round(j["ms_per_step"], 2), "kw_search ms (isolated)", j.get("device_ms_isolated", {}).get("kw_search"),
              "config.kw_scoring", j["config"].get("kw_scoring"))
    except Exception as e:
        print(n, "unreadable:", e)
PY
TSGPU_REG_SCORE=1 ncu --set full --clock-control none --import-source on -k regex:kw_search_kernel -s 1 -c 1 -o $OUT/optin_prof_kw_regscore -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2> $OUT/optin_ncu.log
# 4. f-1: the candidate walk on the device, both forms, next to profiles/r01d_art_cpu.json
python tools/bench_art_gpu.py > $OUT/optin_art_dfs.json 2> $OUT/optin_art_dfs.err
TSGPU_ART_MODE=frontier python tools/bench_art_gpu.py > $OUT/optin_art_frontier.json 2> $OUT/optin_art_frontier.err
tail -1 $OUT/optin_art_dfs.json; tail -1 $OUT/optin_art_frontier.json
ls -la $OUT | grep optin
