import json, sys
d = json.load(open(sys.argv[1]))
for k in ("value", "ms_per_step", "e2e", "device_ms_per_step", "device_ms_isolated", "cpu_baseline", "parity_sample", "knn_recall_at_100", "gpu_launches", "clocks"):
    print(k, json.dumps(d.get(k))[:400])
print("roofline", json.dumps(d.get("roofline"))[:500])
print("roofline_other", json.dumps(d.get("roofline_other"))[:500])
