#!/bin/bash
# Round 2, call 37: cost-volume kernel with h1 as a TMEM A operand of GEMM2: parity, timeline, timing
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_cost_volume.py -q -x > $O/s2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -12 $O/s2_pytest.txt
DTB200_DEVELOPMENT=1 DTB200_CONV_FLAGS=4096 timeout 200 python tools/cv_bench.py --math tch --reps 2 2>&1 | tail -6 | cut -c1-400
timeout 200 python tools/cv_bench.py --math tch --reps 10
