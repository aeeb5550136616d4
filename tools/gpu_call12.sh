#!/bin/bash
# halo-tile conv kernel: parity with the switch on, per-layer A/B, bench A/B
O=gpurun_out
mkdir -p $O
DTB200_CONV_FLAGS=128 timeout 300 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py tests/test_gpu_full_size.py -x -q > $O/g_pytest_halo.txt 2>&1; echo "pytest halo rc=$?"; tail -25 $O/g_pytest_halo.txt
timeout 120 python tools/conv_bench.py --debug 0,128 > $O/g_conv_bench.txt 2>&1; cat $O/g_conv_bench.txt
DTB200_CONV_FLAGS=128 timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/g_bench_halo.json 2> $O/g_bench_halo.err; echo "bench rc=$?"; cut -c1-200 $O/g_bench_halo.json; tail -3 $O/g_bench_halo.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/g_bench_halo.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "conv", d["roofline"]["ms_all_launches"], "cv", d["roofline_kernels"]["cost_volume_mlp_hint"]["ms_per_launch"])
except Exception as e:
    print("parse failed", e)
PY
