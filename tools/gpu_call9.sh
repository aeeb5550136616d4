#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py -q 2>&1 | tail -4
timeout 200 python tools/conv_bench.py --debug 0,16384 2>&1 | grep -v "^total" > $O/c9_conv_bench.txt; cat $O/c9_conv_bench.txt
timeout 200 python tools/conv_bench.py --only s0_64_64_3x3 --reps 1 --debug 8192,24576 > $O/c9_roles.txt 2>&1
python - <<'PY'
txt=open("gpurun_out/c9_roles.txt").read()
for block in txt.split("debug flags")[1:]:
    lines=[l for l in block.splitlines() if l.startswith("role") or l.startswith("latency")]
    n=len(lines)//4 if len(lines)>=4 else len(lines)
    print("debug flags", block.splitlines()[0]); print("\n".join(sorted(lines[-n:])))
PY
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-250
