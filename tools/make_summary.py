#!/usr/bin/env python
"""profiles/r02_summary.md + profiles/traffic.json from the round's committed artefacts:
    tools/make_summary.py [tag]      (tag: the file prefix of the final run, default r02z)
Reads profiles/<tag>_bench.json, <tag>_bench_reference.json, <tag>_ncu_full.md, <tag>_launches.csv, <tag>_bench_*gpu.json (else r02v_bench_*gpu.json)."""
import collections
import csv
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def last_json(path):
    with open(path) as f:
        lines = [l for l in f.read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


def ncu_sections(path):
    out, cur = {}, None
    for line in open(path):
        m = re.match(r"## (\S+) — (?:void )?(?:\w+::)*(\w+)", line)
        if m:
            cur = m.group(2)
            out[cur] = {}
            continue
        m = re.match(r"\| ([^|]+) \| ([^|]+) \| ([^|]*) \|", line)
        if m and cur and m.group(1).strip() != "metric" and not m.group(1).startswith("---"):
            try:
                out[cur][m.group(1).strip()] = (float(m.group(2)), m.group(3).strip())
            except ValueError:
                pass
    return out


def launch_shares(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) != len(hdr):
            continue
        k = r[ix["Kernel Name"]].split("(")[0]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else (v if u.startswith("m") else v * 1e3))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    return agg


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02z"
    b = last_json(os.path.join(P, f"{tag}_bench.json"))
    L = [f"# Round 2: where the bench line's numbers come from (`profiles/{tag}_bench.json`, one B200)", ""]
    roofs = [b["roofline"]] + b.get("roofline_other", [])
    L += [f"* `value` **{b['value']:.0f} {b['unit']}** ({b['ms_per_step']:.2f} ms per 4096-query step: resolved combinations, buffers resident in HBM, one `tsgpu_hybrid_search_batch`).",
          f"* `e2e` **{b['e2e']['value']:.0f} {b['unit']}** ({b['e2e']['ms_per_step']:.1f} ms per step; query strings in host memory -> C++ host layer -> device rounds -> KV records in host memory; "
          f"{b['e2e'].get('calls_in_flight')} requests in flight; h2d {b['e2e']['h2d_bytes_per_step'] / 1e6:.1f} MB, d2h {b['e2e']['d2h_bytes_per_step'] / 1e6:.1f} MB per step, "
          f"{b['e2e'].get('device_calls_per_step', 0):.0f} C-ABI calls per step)."]
    cb = b.get("cpu_baseline")
    if cb:
        L.append(f"* `cpu_baseline` {cb['value']:.1f} {cb['unit']} on {cb['cores']} host threads ({cb['kind']}; {cb['sample']}) -> e2e / cpu = {b['e2e']['value'] / cb['value']:.0f}x.")
    ref_path = os.path.join(P, f"{tag}_bench_reference.json")
    if os.path.exists(ref_path):
        r = last_json(ref_path)
        L.append(f"* `--impl reference` on the same box: {r['value']:.1f} {r['unit']} ({r['ms_per_step']:.0f} ms per step of {r['config'].get('sample_queries', '?')} queries, {r['cpu_baseline']['cores']} threads).")
    ps = b.get("parity_sample")
    if ps:
        L.append(f"* parity sample (GPU e2e arm vs CPU arm, same query strings): {json.dumps(ps)}")
    if "recall" in b:
        L.append(f"* recall: {json.dumps(b['recall'])}")
    if "latency_ms" in b:
        lm = b["latency_ms"]
        L.append(f"* latency of one multi_search through the host layer: 4096 queries p50 {lm['batch']['p50']:.0f} ms / p99 {lm['batch']['p99']:.0f} ms; 64 queries p50 {lm['small']['p50']:.1f} ms / p99 {lm['small']['p99']:.1f} ms.")
    L.append(f"* clocks during the timed regions: {json.dumps(b.get('clocks'))}")
    L += ["", "## Roofline inputs (kernel timed alone, CUDA events on the library's stream, same batches)", "",
          "| kernel | ms | algorithmic GB per launch | achieved GB/s | peak GB/s | frac | DRAM traffic GB (ncu) |", "|---|---|---|---|---|---|---|"]
    ncu_path = os.path.join(P, f"{tag}_ncu_full.md")
    sec = ncu_sections(ncu_path) if os.path.exists(ncu_path) else {}
    traffic = {}
    for k, d in sec.items():
        t = d.get("dram traffic (read+write)")
        if t:
            traffic[k] = t[0] * (1e9 if t[1].startswith("G") else 1e6)
    for r in roofs:
        t = traffic.get(r["kernel"])
        L.append(f"| `{r['kernel']}` | {r['ms']:.2f} | {r['algorithmic_bytes'] / 1e9:.2f} | {r['achieved']:.0f} | {r['peak']:.0f} | **{r['frac']:.3f}** | {t / 1e9:.2f} |" if t else
                 f"| `{r['kernel']}` | {r['ms']:.2f} | {r['algorithmic_bytes'] / 1e9:.2f} | {r['achieved']:.0f} | {r['peak']:.0f} | **{r['frac']:.3f}** | – |")
    L += ["", f"`peak` = {roofs[0].get('peak_source')}. `traffic` = dram__bytes_read.sum + dram__bytes_write.sum of ONE `ncu --set full` launch of this build (`{tag}_ncu_full.md`; "
          "under ncu the kernel runs cold-cache and alone).", ""]
    if traffic:
        traffic["source"] = f"profiles/{tag}_ncu_full.md (ncu --set full, one launch each of this build: dram__bytes_read.sum + dram__bytes_write.sum)"
        json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    w = b.get("work_per_step", {})
    if w:
        L.append(f"Work per step (kernel counters): {json.dumps({k: round(v) for k, v in w.items()})}; walks: {json.dumps(b.get('knn_walks'))}.")
    if "host_rounds_per_step" in b:
        L.append(f"Host layer per 4096-query request: {json.dumps({k: round(v, 1) for k, v in b['host_rounds_per_step'].items()})}.")
    L.append(f"Device time per step, value leg (stream overlap on): {json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in b.get('device_ms_per_step', {}).items()})}; "
             f"kernels alone: {json.dumps({k: round(v, 2) for k, v in b.get('device_ms_isolated', {}).items()})}.")
    lpath = os.path.join(P, f"{tag}_launches.csv")
    if os.path.exists(lpath):
        agg = launch_shares(lpath)
        tot = sum(a[1] for a in agg.values())
        ours = {k: a for k, a in agg.items() if re.search(r"ts[a-z_]*::|art_|hnsw_|kw_|facet|filter|setop|isect", k)}
        L += ["", f"## Kernel shares of the whole bench command under ncu (`{tag}_launches.csv`: {sum(a[0] for a in agg.values())} launches, "
              f"{tot:.0f} ms of kernel time, cold-cache and serialised; torch kernels = synthetic data generation of the harness)", "",
              "| kernel | launches | total ms | share of the library's kernels |", "|---|---|---|---|"]
        tot_ours = sum(a[1] for a in ours.values())
        for k, a in sorted(ours.items(), key=lambda x: -x[1][1])[:12]:
            L.append(f"| `{k[:80]}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot_ours:.1f} % |")
        # the resolved-query legs (value, e2e_resolved): from the first kw_search launch to the first ART walk launch
        rows = list(csv.reader(open(lpath)))
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        ixk = rows[hi].index("Kernel Name"); ixv = rows[hi].index("Metric Value"); ixu = rows[hi].index("Metric Unit")
        win, started = collections.OrderedDict(), False
        for r in rows[hi + 1:]:
            if len(r) != len(rows[hi]):
                continue
            k = r[ixk].split("(")[0]
            if "art_frontier" in k:
                break
            if "kw_search_kernel" in k:
                started = True
            if not started or not re.search(r"ts[a-z_]*::", k):
                continue
            v = float(r[ixv].replace(",", "")); u = r[ixu]
            ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
            a = win.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += ms
        if win:
            tw = sum(a[1] for a in win.values())
            L += ["", "Resolved-query legs only (launches before the first ART walk), per-kernel share of the step under ncu vs. live:", "",
                  "| kernel | launches | mean ms under ncu | share under ncu | share live (kernels alone) |", "|---|---|---|---|---|"]
            iso = b.get("device_ms_isolated", {})
            live = {"hnsw_walk_kernel": iso.get("knn", 0), "kw_search_kernel": iso.get("kw_search", 0), "kw_merge_kernel": iso.get("kw_merge", 0)}
            tl = iso.get("total", 0) or 1
            for k, a in sorted(win.items(), key=lambda x: -x[1][1])[:6]:
                name = next((n for n in live if n in k), None)
                real = [x for x in [a[1] / max(1, a[0])]]
                L.append(f"| `{k[:70]}` | {a[0]} | {a[1] / a[0]:.3f} | {100 * a[1] / tw:.1f} % | {('%.1f %%' % (100 * live[name] / tl)) if name else '–'} |")
            L.append("")
            L.append("(every hybrid call launches the walk kernel twice: the main launch and the retry launch, which ends at once when no walk overflowed — the mean above averages both.)")
        L.append("")
        L.append("In the value leg the walk kernel and kw_search split the step as the live CUDA-event times above do "
                 f"({b['device_ms_isolated'].get('knn', 0):.1f} : {b['device_ms_isolated'].get('kw_search', 0):.1f} ms); the ART frontier launches belong to the end-to-end leg.")
    sc = []
    scale_files = glob.glob(os.path.join(P, f"{tag}_bench_*gpu.json")) or glob.glob(os.path.join(P, "r02v_bench_*gpu.json"))      # this build's, else the round's earlier ones
    for f in sorted(scale_files, key=lambda x: int(re.search(r"_(\d+)gpu", x).group(1))):
        sc.append(last_json(f))
    if sc:
        L += ["", "## Strong scaling (one 4096-query request, every rank answers its slice, gather inside the library)", "",
              "| GPUs | value q/s | ms per step | e2e q/s | gather ms (median, rank 0) | walk kernel ms (slice) |", "|---|---|---|---|---|---|",
              f"| 1 | {b['value']:.0f} | {b['ms_per_step']:.2f} | {b['e2e']['value']:.0f} | – | {b['device_ms_isolated'].get('knn', 0):.2f} |"]
        for j in sc:
            c = j.get("collective", {})
            L.append(f"| {j['n_gpus']} | {j['value']:.0f} | {j['ms_per_step']:.2f} | {j['e2e']['value']:.0f} | {c.get('ms_median_rank0', 0):.2f} | {j['device_ms_isolated'].get('knn', 0):.2f} |")
        L.append("")
        L.append("The slice's walk kernel does not shrink with the slice: its time is the longest walk's latency (DESIGN.md §4.1), which is what bounds strong scaling here.")
    oc = b.get("other_configs")
    if oc:
        L += ["", "## Secondary configurations (same JSON line, `other_configs`)", "", "```", json.dumps(oc, indent=1)[:6000], "```"]
    open(os.path.join(P, "r02_summary.md"), "w").write("\n".join(L) + "\n")
    print("\n".join(L[:40]))


if __name__ == "__main__":
    main()
