#!/bin/bash
# last call of the round: canonical bench line of HEAD (halo kernel, 3 taps per weight stage by default) + halo parity tests
O=gpurun_out
mkdir -p $O
timeout 150 python bench.py --steps 20 --warmup 3 --cpu-budget 12 > $O/i_bench_tc3x.json 2> $O/i_bench_tc3x.err; echo "bench rc=$?"; cut -c1-200 $O/i_bench_tc3x.json; tail -3 $O/i_bench_tc3x.err
timeout 100 python -m pytest tests/test_gpu_networks.py tests/test_gpu_full_size.py -x -q > $O/i_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/i_pytest.txt
