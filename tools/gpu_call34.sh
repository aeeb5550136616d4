#!/bin/bash
# Round 2, call 20: unrolled compile-time MMA issue sequences (all three tch conv kernels): parity, per-layer timing, trace, bench
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py tests/test_gpu_full_size.py -q -x > $O/d2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/d2_pytest.txt
timeout 100 python tools/conv_bench.py --math tch --only s0_64_64_3x3 --reps 2 --debug 4096 2>&1 | tail -8 | cut -c1-330
timeout 300 python tools/conv_bench.py --math tch --debug 0,16 > $O/d2_conv_bench.txt 2>&1; cat $O/d2_conv_bench.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/d2_bench.json 2> $O/d2_bench.err; echo "bench rc=$?"; cat $O/d2_bench.json
