#!/bin/bash
# Round 2, call 10: conv tch with 8 epilogue warps.
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py -q -x > $O/t_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/t_pytest.txt
timeout 100 python tools/conv_bench.py --math tch > $O/t_conv_bench.txt 2>&1; tail -15 $O/t_conv_bench.txt
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu --sustain 0 > $O/t_bench.json 2> $O/t_bench.err; echo "bench rc=$?"; cut -c1-200 $O/t_bench.json; tail -2 $O/t_bench.err
