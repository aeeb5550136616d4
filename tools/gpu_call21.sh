#!/bin/bash
# Round 2, call 6: first run of the fp16-split conv kernels (math = tch, split16 activations).
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_networks.py -q -x -k "tch" > $O/p_pytest_net.txt 2>&1; echo "pytest networks tch rc=$?"; tail -25 $O/p_pytest_net.txt
timeout 300 python -m pytest tests/test_gpu_model.py -q -x -k "tch" > $O/p_pytest_model.txt 2>&1; echo "pytest model tch rc=$?"; tail -5 $O/p_pytest_model.txt
timeout 100 python tools/conv_bench.py --math tch > $O/p_conv_bench.txt 2>&1; tail -15 $O/p_conv_bench.txt
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/p_bench.json 2> $O/p_bench.err; echo "bench rc=$?"; cut -c1-200 $O/p_bench.json; tail -3 $O/p_bench.err
