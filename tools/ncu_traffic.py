"""Sum the per-launch DRAM bytes and durations of one profiled step (csv written by ncu around tools/profile_step.py) and
merge them into profiles/traffic.json under <workload>_<dtype>; prints the per-kernel table kept under profiles/."""
import csv, json, os, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, key = sys.argv[1], sys.argv[2]
rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
agg = defaultdict(lambda: defaultdict(float))
cnt = defaultdict(set)
for r in rows:
    k = r["Kernel Name"].split("(")[0].split("::")[-1][:60]
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(u, 1)
    agg[k][r["Metric Name"]] += v * scale
    cnt[k].add(r["ID"])
tot_b = tot_ms = 0.0
print(f"# {key}: one step, per kernel (ncu, serialised launches: compare shares)")
for k, m in agg.items():
    b = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
    ms = m.get("gpu__time_duration.sum", 0)
    tot_b += b; tot_ms += ms
    print(f"{k:60s} launches {len(cnt[k]):3d}  time {ms:8.3f} ms  dram read {m.get('dram__bytes_read.sum', 0) / 1e9:7.3f} GB  write {m.get('dram__bytes_write.sum', 0) / 1e9:7.3f} GB")
print(f"{'total':60s}               time {tot_ms:8.3f} ms  dram {tot_b / 1e9:7.3f} GB")
tp = os.path.join(ROOT, "profiles", "traffic.json")
d = json.load(open(tp)) if os.path.exists(tp) else {}
d[key] = tot_b
json.dump(d, open(tp, "w"), indent=1, sort_keys=True)
