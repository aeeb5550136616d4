#!/bin/bash
O=gpurun_out
mkdir -p $O
# per-role accounting (32<<8 = 8192): full, skeleton, MMA-only, split-only
K="8192,16128,11776,11520"
timeout 200 python tools/conv_bench.py --only s0_64_64_3x3 --reps 1 --debug $K > $O/c6_roles.txt 2>&1
python - <<'PY'
import re,collections
txt=open("gpurun_out/c6_roles.txt").read()
# keep the last launch's report of each variant (4 launches per variant: 3 warm + 1 timed)
for block in txt.split("debug flags")[1:]:
    lines=[l for l in block.splitlines() if l.startswith("role")]
    n=len(lines)//4 if len(lines)>=4 else len(lines)
    print("debug flags", block.splitlines()[0]); print("\n".join(lines[-n:]))
    print([l for l in block.splitlines() if "TFLOP" in l])
PY
