"""The measurement tools under tools/ keep working: the CPU timing of the candidate search, and (through the test double) the
Python side of its GPU counterpart."""
import json
import os
import subprocess
import sys

import pytest

import oracle_lib as ol

ROOT = ol.ROOT


@pytest.mark.skipif(not ol.have_ref() or not hasattr(ol.ref(), "ref_art_new"), reason="oracle/_ref with the reference's art.cpp not built")
def test_bench_art_cpu_runs():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_art_cpu.py"), "3000"], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads(r.stdout)
    assert j["tokens"] > 2000 and set(j["per_search_us"]) == {"prefix0", "typo1", "typo2"}
    assert j["per_search_us"]["typo2"]["nodes_visited_per_search"] > j["per_search_us"]["prefix0"]["nodes_visited_per_search"]


def test_bench_art_gpu_python_side_on_the_double():
    so = os.path.join(ROOT, "tests", "cpp", "libtsgpu_double.so")
    if not os.path.exists(so):
        import test_gpu_tests_dryrun  # noqa: F401  (its first test builds the double)
        pytest.skip("double not built yet")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_art_gpu.py"), "3000", "64"], capture_output=True, text=True, cwd=ROOT, timeout=300,
                       env=dict(os.environ, TSGPU_TEST_DOUBLE="1", TSGPU_LIB_PATH=so))
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads(r.stdout.strip().splitlines()[-1])
    assert all(k["sample_mismatches"] == 0 for k in j["kinds"].values())
