#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 75 python -m pytest tests/test_gpu_tsdf.py -x -q > $O/j_pytest_tsdf.txt 2>&1; echo "pytest tsdf rc=$?"; tail -12 $O/j_pytest_tsdf.txt
timeout 40 python tools/tsdf_bench.py --reps 5 > $O/j_tsdf_bench.txt 2>&1; echo "bench rc=$?"; cat $O/j_tsdf_bench.txt | tail -8
