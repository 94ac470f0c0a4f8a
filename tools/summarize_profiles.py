"""Turn the ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/:
   launch lists  -> profiles/<name>_launches_summary.txt (time share per kernel)
   .ncu-rep      -> profiles/<name>_ncu_summary.txt (key counters + top stall lines), profiles/traffic.json"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

def launches(csv_path, out_name):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10]
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    rec = {}
    for r in rows[1:]:
        key = (r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0])
        rec.setdefault(key, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    tot = collections.defaultdict(float); cnt = collections.Counter()
    for (i, k), m in rec.items():
        tot[k] += m.get("gpu__time_duration.sum", 0); cnt[k] += 1
    T = sum(tot.values())
    with open(os.path.join(OUT, out_name), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# source: %s, %d launches, %.3f ms total\n" % (os.path.basename(csv_path), len(rec), T / 1e6))
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write("%-28s launches=%5d  time=%10.1f us  share=%5.1f%%\n" % (k, cnt[k], v / 1e3, 100 * v / T))
    print(open(os.path.join(OUT, out_name)).read())

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max"]

def report(rep, out_name, traffic_key=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    lines = ["# ncu --set full --clock-control none --import-source on; report %s" % os.path.basename(rep)]
    traffic = None
    per_kernel = {}
    for row in rows[2:]:
        lines.append("kernel: " + row[h.index("Kernel Name")][:100])
        vals = {}
        for k in KEYS:
            if k in h:
                vals[k] = row[h.index(k)]
                lines.append("  %-75s %s %s" % (k, row[h.index(k)], units[h.index(k)]))
        try:
            def tobytes(k):
                v = float(vals[k].replace(",", "")); u = units[h.index(k)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            traffic = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
            lines.append("  dram traffic per launch (read+write): %.1f MB" % (traffic / 1e6))
            per_kernel.setdefault(row[h.index("Kernel Name")].split("(")[0], traffic)
        except Exception:
            pass
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Line No"]
    if hi:
        hdr = rows[hi[0]]; end = hi[1] if len(hi) > 1 else len(rows)
        ix = {}
        for i, hh in enumerate(hdr): ix.setdefault(hh, i)
        def f(r, k):
            try: return float(r[ix[k]])
            except Exception: return 0.0
        agg = [r for r in rows[hi[0] + 1:end] if len(r) > 10 and r[2] == "-" and r[0].isdigit()]
        tot = sum(f(r, "# Samples") for r in agg) or 1
        st = {k: sum(f(r, k) for r in agg) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
        lines.append("stall samples (first profiled launch): " + ", ".join("%s %d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
        lines.append("top source lines by samples:")
        for r in sorted(agg, key=lambda r: -f(r, "# Samples"))[:12]:
            s2 = {k: f(r, k) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
            big = sorted(s2.items(), key=lambda kv: -kv[1])[:2]
            lines.append("  %5.1f%%  L%-4s %-90s %s" % (100 * f(r, "# Samples") / tot, r[0], r[1][:90].strip(), [(k[6:], int(v)) for k, v in big]))
    open(os.path.join(OUT, out_name), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))
    if traffic_key and traffic is not None:
        tp = os.path.join(OUT, "traffic.json")
        t = json.load(open(tp)) if os.path.exists(tp) else {}
        if isinstance(traffic_key, dict):          # kernel name -> key, first launch of each kernel
            for kn, key in traffic_key.items():
                if kn in per_kernel: t[key] = per_kernel[kn]
        else:
            t[traffic_key] = traffic
        json.dump(t, open(tp, "w"), indent=1)

if __name__ == "__main__":
    g = os.path.join(ROOT, "gpurun_out")
    for csvf, name in (("r01_launches_bench_klu.csv", "r01_launches_bench_klu_summary.txt"), ("r01_launches_chol64.csv", "r01_launches_chol64_summary.txt")):
        if os.path.exists(os.path.join(g, csvf)):
            launches(os.path.join(g, csvf), name)
    for rep, name, key in (("r01_klu_wave.ncu-rep", "r01_ncu_k_klu_refactor_wave.txt", "k_klu_refactor"), ("r01_chol_update.ncu-rep", "r01_ncu_k_update.txt", "k_update"),
                           ("r01_klu_refactor_final.ncu-rep", "r01_ncu_klu_refactor_final.txt",
                            {"k_klu_refactor_wave": "k_klu_refactor_wave", "k_klu_dense_lu": "k_klu_dense_lu", "k_klu_dense_pack": "k_klu_dense_pack"})):
        if os.path.exists(os.path.join(g, rep)):
            report(os.path.join(g, rep), name, key)
