"""The register-resident scoring kernel (kw_regscore.cu) is the default since round 2 (29.4 -> 19.4 ms per 4096-query batch),
so every keyword GPU test runs it. This file keeps the round-1 kernel (kw_search_kernel<false>, TSGPU_REG_SCORE=0: the
fallback for A/B runs) honest: the switch is read once per process, so the keyword parity tests and the reference
scenarios are re-run in a child process with it set."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_keyword_parity_with_r01_scoring_kernel():
    env = dict(os.environ, TSGPU_REG_SCORE="0")
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
           os.path.join(ROOT, "tests", "test_gpu_parity.py"), os.path.join(ROOT, "tests", "test_reference_scenarios.py"),
           os.path.join(ROOT, "tests", "test_typo_scenarios.py"), os.path.join(ROOT, "tests", "test_specific_scenarios.py"),
           "-k", "keyword or scenarios or edge or large_scale"]
    if os.environ.get("TSGPU_TEST_DOUBLE") == "1":            # CPU dry run (tests/test_gpu_tests_dryrun.py): leave out what needs a device
        from test_gpu_tests_dryrun import GPU_ONLY
        for t in GPU_ONLY:
            cmd += ["--deselect", t]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=480)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and " passed" in r.stdout, tail
