"""profiles/r02_traffic.json from an ncu launch list (tools/summarize_ncu.py's input): measured DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum) of the LAST training step in the capture, per bench.py kernel class.
Conv launches are split by grid size: streamed-weight (>= 128-channel) launches use fewer CTAs than SMs at configs[1]."""
import collections, csv, gzip, json, re, sys

src, out, workload = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "train_base_b16")
op = gzip.open if src.endswith(".gz") else open
rows = list(csv.DictReader([l for l in op(src, "rt") if l.startswith('"')]))
per = collections.OrderedDict()
for r in rows:
    d = per.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]), "grid": int(r["Grid Size"].strip("()").split(",")[0])})
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
    if r["Metric Name"].endswith("bytes_read.sum") or r["Metric Name"].endswith("bytes_write.sum"):
        d["dram"] = d.get("dram", 0.0) + v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
ids = sorted(per)
packs = [i for i in ids if "wn_pack" in per[i]["name"]]
ids = [i for i in ids if i >= packs[-1]]          # the last step starts with its weight-norm fold
note = (f"{src} (ncu dram__bytes_read.sum + dram__bytes_write.sum, last training step of the capture, cold caches; conv launches "
        "split by grid < 148 CTAs = streamed-weight >= 128-channel layers)")
cls = collections.OrderedDict()
for i in ids:
    d = per[i]
    n = d["name"]
    if "wgrad_kernel" in n:
        key = "tc_wgrad_c>=128"            # bench.py reads the all-wgrad average for both wgrad classes
    elif "pair_kernel" in n or ("conv_kernel" in n and d["grid"] >= 148):
        key = "tc_conv_c<=64(fwd+dgrad,pair)"
    elif "conv_kernel" in n:
        key = "tc_conv_c>=128(fwd+dgrad)"
    elif "conv_post" in n:
        key = "conv_post"
    elif "wn_" in n:
        key = "weight_norm_fold"
    else:
        continue
    a = cls.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += d.get("dram", 0.0)
res = {workload: {k: {"dram_bytes_per_launch": v[1] / v[0], "launches": v[0],
                      "source": note + (" (all wgrad launches)" if k.startswith("tc_wgrad") else "")} for k, v in cls.items()}}
if "tc_wgrad_c>=128" in res[workload]:
    res[workload]["tc_wgrad_c<=64"] = dict(res[workload]["tc_wgrad_c>=128"])
try:
    old = json.load(open(out))
except Exception:
    old = {}
old.update(res)
json.dump(old, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
