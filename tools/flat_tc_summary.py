#!/usr/bin/env python
"""profiles/flat_tc.json from the check program's log (tests/cpp/flat_tc_check.cpp) and the ncu summary of flat_tc_kernel:
    python tools/flat_tc_summary.py gpurun_out/flat_tc_check.log profiles/<ncu summary>.md"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = open(sys.argv[1]).read()
rows = []
for m in re.finditer(r"(ok|FAIL)\s+nq=(\d+) ids=(\d+) dim=(\d+): tc_queries=(\d+) max_abs=(\S+) max_rel\(\|ref\|>1e-3\)=(\S+) vs_scalar=(\S+) bad=(\d+) kernel_ms=(\S+) \((\S+) TFLOP/s fp32-equivalent, (\S+) GB/s", log):
    rows.append(dict(status=m[1], nq=int(m[2]), ids=int(m[3]), dim=int(m[4]), max_abs=float(m[6]), max_rel=float(m[7]), vs_scalar=float(m[8]), bad=int(m[9]),
                     kernel_ms=float(m[10]), tflops_fp32_equiv=float(m[11]), rows_gbs=float(m[12])))
big = [r for r in rows if r["ids"] >= 100000]
pipe = None
if len(sys.argv) > 2:
    txt = open(sys.argv[2]).read()
    i = txt.find("flat_tc_kernel")                      # the summary may hold several kernels: this kernel's section
    m = re.search(r"sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \| ([\d.]+)", txt[i if i >= 0 else 0:])
    pipe = round(float(m[1]), 1) if m else None
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
best = {}
for r in big:
    if r["nq"] not in best or r["kernel_ms"] < best[r["nq"]]["kernel_ms"]:
        best[r["nq"]] = r
perf = "; ".join(f"{nq} queries: {b['kernel_ms']:.3f} ms = {b['tflops_fp32_equiv']:.0f} TFLOP/s fp32-equivalent ({3 * b['tflops_fp32_equiv']:.0f} TFLOP/s of tf32 MMAs), rows at {b['rows_gbs']:.0f} GB/s"
                 for nq, b in sorted(best.items(), reverse=True))
dev768 = max((r["max_abs"] for r in rows if r["dim"] == 768), default=None)
out = {"cases": rows, "tensor_pipe_pct": pipe, "perf": perf, "ncu_file": os.path.basename(sys.argv[2]) if len(sys.argv) > 2 else None,
       "deviation": f"largest deviation from a double-precision dot over all cases of the check: {max(r['max_abs'] for r in rows):.2g} absolute ({dev768:.2g} at d = 768)",
       "hbm_peak_gbs": peaks.get("hbm_gbs"), "bf16_peak_tflops": peaks.get("bf16_tflops"),
       "note": "tf32 dense peak is nominally half the bf16 one (B200_PROFILING.md: 1.1 PFLOP/s); tensor roofline frac of the 256-query case = 3 x fp32-equivalent TFLOP/s / (MEASURED_PEAKS bf16_tflops / 2)"}
if 256 in best:
    out["tensor_roofline_frac_256q"] = round(3 * best[256]["tflops_fp32_equiv"] / ((peaks.get("bf16_tflops") or 2200.0) / 2), 3)
if 16 in best and peaks.get("hbm_gbs"):
    out["hbm_roofline_frac_16q"] = round(best[16]["rows_gbs"] / peaks["hbm_gbs"], 3)
json.dump(out, open(os.path.join(ROOT, "profiles", "flat_tc.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in out.items() if k != "cases"}, indent=1))
