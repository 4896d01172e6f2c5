"""Summarise an ncu launch list (CSV of gpu__time_duration.sum) and a `--set full` report (raw page CSV)
into the tables committed under profiles/.   python tools/ncu_summarise.py <launches.csv> <raw.csv> <tag>"""
import collections, csv, re, sys

launches, raw, tag = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(launches) if not l.startswith("==")]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print("| kernel | launches | avg us | share |\n|---|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if t / tot >= 0.002:
        print(f"| `{k}` | {n} | {t / n:.1f} | {t / tot:.1%} |")
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
idx = [hdr.index(w) if w in hdr else None for w in want]
with open(f"profiles/{tag}_ncu_full_metrics.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(want); w.writerow([units[i] if i is not None else "" for i in idx])
    for r in rows[2:]:
        w.writerow([(r[i] if i is not None else "") for i in idx])
        print([(r[i][:48] if i is not None else "") for i in idx[:11]])
