"""ncu per-launch CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) -> per-step summary.

    python tools/traffic_from_launches.py gpurun_out/launches.csv tc3x steps_profiled > profiles/…
Prints a JSON fragment {family: dram bytes per step} and a text table of time shares per kernel."""
import collections
import csv
import json
import sys

path, math, steps = sys.argv[1], sys.argv[2], float(sys.argv[3])
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
idx = {n: i for i, n in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    per.setdefault((r[idx["ID"]], r[idx["Kernel Name"]]), {})[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
fam = lambda k: ("cost_volume" if "cv_" in k else "conv_stack" if ("conv_" in k or "splitk" in k or "resample" in k) else "other")
t = collections.defaultdict(float)
byt = collections.defaultdict(float)
cnt = collections.Counter()
tk = collections.defaultdict(float)
ck = collections.Counter()
for (i, k), m in per.items():
    f = fam(k)
    dur = m.get("gpu__time_duration.sum", 0.0)
    t[f] += dur
    byt[f] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    cnt[f] += 1
    name = k.split("(")[0][-40:]
    tk[name] += dur
    ck[name] += 1
unit = 1e6 if max(byt.values()) < 1e7 else 1.0  # ncu prints Mbyte
total = sum(t.values())
print(json.dumps({math: {f: int(byt[f] * unit / steps) for f in byt}}))
print(f"# {steps} steps profiled; time per step and share (ncu serialised, cold caches: compare shares, not absolutes)", file=sys.stderr)
for name, v in sorted(tk.items(), key=lambda kv: -kv[1]):
    print(f"# {name:42s} launches/step {ck[name] / steps:6.1f}  time/step {v / steps / 1e3:9.1f} us  share {v / total * 100:5.1f} %", file=sys.stderr)
