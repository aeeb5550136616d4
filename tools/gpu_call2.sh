#!/bin/bash
# gpurun call: validate graph mode + merged-N MMA; A/B benches.
O=gpurun_out
mkdir -p $O
timeout 500 python -m pytest tests -m gpu -q > $O/c2_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -15 $O/c2_pytest_gpu.txt
timeout 300 python bench.py --steps 20 --warmup 3 > $O/c2_bench_graph.json 2> $O/c2_bench_graph.err; echo "bench rc=$?"; cut -c1-260 $O/c2_bench_graph.json; tail -3 $O/c2_bench_graph.err
for v in "MODE=sequence" "LANES=1" "LANES=4" "LANES=16" "WS_SLOTS=1"; do
  env DTB200_CONV_$v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/c2_bench_$v.json 2> $O/c2_bench_$v.err
  echo "== $v rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$O/c2_bench_$v.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["ms_all_launches"], d["gpu_launches"])
except Exception as e:
    print("parse failed", e)
PY
done
for f in 0 1 2; do echo "== conv_bench flags $f"; timeout 150 python tools/conv_bench.py --debug $f 2>&1 | tail -15; done > $O/c2_conv_bench.txt 2>&1; cat $O/c2_conv_bench.txt
DTB200_CONV_FLAGS=2 timeout 300 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py -q > $O/c2_pytest_flags2.txt 2>&1; echo "flags2 pytest rc=$?"; tail -5 $O/c2_pytest_flags2.txt
DTB200_CONV_FLAGS=2 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/c2_bench_flags2.json 2>&1; cut -c1-200 $O/c2_bench_flags2.json
DTB200_CONV_FLAGS=1 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/c2_bench_flags1.json 2>&1; cut -c1-200 $O/c2_bench_flags1.json
