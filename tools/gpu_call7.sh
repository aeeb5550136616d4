#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py -q 2>&1 | tail -4
timeout 200 python tools/conv_bench.py 2>&1 | tail -15 > $O/c7_conv_bench.txt; cat $O/c7_conv_bench.txt
timeout 200 python tools/conv_bench.py --only s0_64_64_3x3 --reps 1 --debug 8192,16128 > $O/c7_roles.txt 2>&1
python - <<'PY'
txt=open("gpurun_out/c7_roles.txt").read()
for block in txt.split("debug flags")[1:]:
    lines=[l for l in block.splitlines() if l.startswith("role")]
    n=len(lines)//4 if len(lines)>=4 else len(lines)
    print("debug flags", block.splitlines()[0]); print("\n".join(sorted(lines[-n:])))
PY
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-250
