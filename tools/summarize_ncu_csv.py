"""Aggregate an `ncu --csv` metric log (long format: one row per launch x metric) by kernel.

usage: summarize_ncu_csv.py launches.csv [--json out.json --match tc_gemm_kernel]
Prints per-kernel launch count, total / average gpu__time_duration and (when captured) DRAM bytes; with --json writes
{"launches", "dram_bytes_total", "time_us"} for the kernels whose name contains --match."""
import csv, collections, json, re, sys
path = sys.argv[1]
match = sys.argv[sys.argv.index("--match") + 1] if "--match" in sys.argv else None
out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
lines = [l for l in open(path, errors="replace") if l.startswith('"')]
rows = list(csv.DictReader(lines))
per_launch = collections.OrderedDict()
for r in rows:
    d = per_launch.setdefault(r["ID"], {"name": r["Kernel Name"]})
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = r["Metric Unit"].lower()
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "second": 1e6, "nsecond": 1e-3,
             "byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
    d[r["Metric Name"]] = v * scale
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in per_launch.values():
    name = re.sub(r"\(.*", "", d["name"]).replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    a = agg[name[:90]]
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values()) or 1.0
print("launches %d   total gpu time %.1f us (serialised, cold-cache: shares are meaningful, absolutes are not)" % (len(per_launch), tot))
for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-92s n=%5d total=%9.1f us avg=%8.1f us share=%.3f%s" % (k, n, t, t / n, t / tot, ("  dram=%.1f MB" % (b / 1e6)) if b else ""))
if out_json and match:
    sel = [(n, t, b) for k, (n, t, b) in agg.items() if re.search(match, k)]      # --match is a regular expression
    json.dump({"kernel": match, "launches": sum(x[0] for x in sel), "time_us": sum(x[1] for x in sel),
               "dram_bytes_total": sum(x[2] for x in sel), "all_kernels_time_us": tot,
               "share_of_step": sum(x[1] for x in sel) / tot, "source": path}, open(out_json, "w"), indent=1)
