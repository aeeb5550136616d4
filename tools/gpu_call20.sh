#!/bin/bash
# Round 2, call 5: back-off sleeps in the non-critical mbarrier waits (cost volume tch + conv kernels).
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_cost_volume.py tests/test_gpu_networks.py tests/test_gpu_model.py -q -x > $O/o_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/o_pytest.txt
for m in tc3x tch; do timeout 60 python tools/cv_bench.py --math $m --reps 10; done 2>&1 | tee $O/o_cv_bench.txt
timeout 100 python tools/conv_bench.py > $O/o_conv_bench.txt 2>&1; tail -15 $O/o_conv_bench.txt
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/o_bench.json 2> $O/o_bench.err; echo "bench rc=$?"; cut -c1-200 $O/o_bench.json; tail -2 $O/o_bench.err
