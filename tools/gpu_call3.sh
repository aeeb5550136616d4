#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 500 python -m pytest tests -m gpu -q > $O/c3_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -15 $O/c3_pytest_gpu.txt
run() {  # name, env...
  n=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/c3_bench_$n.json 2> $O/c3_bench_$n.err
  echo "== $n rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$O/c3_bench_$n.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "conv", d["roofline"]["ms_all_launches"], "cv", d["roofline_kernels"]["cost_volume_mlp_hint"]["ms_per_launch"], d["gpu_launches"])
except Exception as e:
    print("parse failed", e); print(open("$O/c3_bench_$n.err").read()[-2000:])
PY
}
run default X=1
run legacy_split DTB200_CONV_FLAGS=4
run slots12 DTB200_CONV_WS_SLOTS=12
run slots24 DTB200_CONV_WS_SLOTS=24
run lanes12_slots12 DTB200_CONV_WS_SLOTS=12 DTB200_CONV_LANES=12
timeout 150 python tools/conv_bench.py 2>&1 | tail -15 > $O/c3_conv_bench.txt; cat $O/c3_conv_bench.txt
