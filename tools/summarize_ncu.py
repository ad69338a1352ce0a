"""Summarise an ncu --csv launch list (one row per launch x metric) into a per-kernel table (markdown)."""
import collections, csv, re, sys

src, out = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # launches to skip (warm-up steps)
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(lines))
per = collections.OrderedDict()
for r in rows:
    d = per.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]), "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
    if r["Metric Name"].endswith("bytes.sum") or r["Metric Name"].endswith("bytes_read.sum") or r["Metric Name"].endswith("bytes_write.sum"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    d[r["Metric Name"]] = v
ids = sorted(per)[skip:]
agg = collections.OrderedDict()
for i in ids:
    d = per[i]
    a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "dram": 0.0, "l2": 0.0, "tensor": 0.0})
    a["n"] += 1
    a["us"] += d.get("gpu__time_duration.sum", 0.0)
    a["dram"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["l2"] += d.get("lts__t_bytes.sum", 0.0)
    a["tensor"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * d.get("gpu__time_duration.sum", 0.0)
tot = sum(a["us"] for a in agg.values())
with open(out, "w") as f:
    f.write(f"launches analysed: {len(ids)} (skipped {skip}); total serialised device time {tot/1000:.3f} ms\n\n")
    f.write("| kernel | launches | total us | share | avg us | DRAM MB/launch | L2 MB/launch | tensor-pipe active % (time-weighted) |\n|---|---|---|---|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        f.write(f"| {k} | {a['n']} | {a['us']:.1f} | {100*a['us']/tot:.1f}% | {a['us']/a['n']:.1f} | {a['dram']/a['n']/1e6:.2f} | "
                f"{a['l2']/a['n']/1e6:.2f} | {a['tensor']/a['us'] if a['us'] else 0:.1f} |\n")
print(open(out).read())
