"""ncu per-launch CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) -> per-step summary.

    python tools/traffic_from_launches.py gpurun_out/launches.csv tch [conv_launches_per_step=169] > profiles/rNN_traffic.json 2> profiles/rNN_launches_summary.txt

bench.py runs whole steps and, for its roofline legs, each kernel family alone, so the families are normalised separately:
cost-volume steps = number of main cost-volume launches, conv-stack steps = conv-family launches / launches per step.
stdout: {math: {family: DRAM bytes per step}} (bench.py reads profiles/r*_traffic.json for `roofline.traffic`);
stderr: per-kernel launches, time and DRAM bytes per step and the share of the serialised kernel time.  ncu serialises launches and
flushes caches between them: compare shares, not absolutes; DRAM bytes are cold-cache (inter-layer activations live in the 126 MB
L2 in the real step)."""
import collections
import csv
import json
import sys

path, math = sys.argv[1], sys.argv[2]
per_step_conv = float(sys.argv[3]) if len(sys.argv) > 3 else 169.0
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
idx = {n: i for i, n in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    d = per.setdefault(r[idx["ID"]], {"k": r[idx["Kernel Name"]]})
    val = float(r[idx["Metric Value"]].replace(",", ""))
    unit = r[idx["Metric Unit"]]
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}.get(unit, 1.0)
    d[r[idx["Metric Name"]]] = val * scale


def fam(k):
    if "cv_" in k:
        return "cost_volume"
    if any(s in k for s in ("conv_", "splitk", "resample")):
        return "conv_stack"
    return "other"


agg = collections.defaultdict(lambda: [0, 0.0, 0.0])   # launches, ns, bytes
for m in per.values():
    name = m["k"].split("(")[0].replace("void ", "").replace("dtb200::", "")
    a = agg[(fam(m["k"]), name)]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
n_cv = sum(a[0] for (f, n), a in agg.items() if f == "cost_volume" and "cv_mlp" in n or "cv_dot" in n) or 1
n_conv = sum(a[0] for (f, n), a in agg.items() if f == "conv_stack") / per_step_conv or 1
steps = {"cost_volume": n_cv, "conv_stack": n_conv}
byt = collections.defaultdict(float)
tim = collections.defaultdict(float)
for (f, n), a in agg.items():
    if f in steps:
        byt[f] += a[2] / steps[f]
        tim[f] += a[1] / steps[f]
print(json.dumps({math: {f: int(v) for f, v in byt.items()}}))
tot = sum(tim.values())
print(f"# per step: cost volume over {n_cv:.0f} launches, conv stack over {n_conv:.1f} plan replays ({per_step_conv:.0f} launches each); ncu serialised, cold caches",
      file=sys.stderr)
for f in steps:
    print(f"# {f:12s} {tim[f] / 1e3:9.1f} us/step  {byt[f] / 1e6:9.1f} MB DRAM/step  share of kernel time {tim[f] / tot * 100:5.1f} %", file=sys.stderr)
for (f, n), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if f in steps:
        print(f"#   {n[:44]:44s} launches/step {a[0] / steps[f]:6.1f}  time/step {a[1] / steps[f] / 1e3:9.1f} us  DRAM/step {a[2] / steps[f] / 1e6:8.1f} MB",
              file=sys.stderr)
