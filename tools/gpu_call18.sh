#!/bin/bash
# Round 2, call 3: tch volume in the bench, ncu of the tch kernel, TSDF CUDA pin with the fp32-index semantics.
O=gpurun_out
mkdir -p $O
timeout 100 python -m pytest tests/test_gpu_tsdf.py -q -s -k "cuda" > $O/m_pytest_tsdf.txt 2>&1; echo "pytest tsdf rc=$?"; grep "tsdf aten_cuda" $O/m_pytest_tsdf.txt | grep -v print; tail -3 $O/m_pytest_tsdf.txt
timeout 200 python bench.py --steps 20 --warmup 3 --cpu-budget 8 > $O/m_bench.json 2> $O/m_bench.err; echo "bench rc=$?"; cut -c1-300 $O/m_bench.json; tail -2 $O/m_bench.err
timeout 120 ncu --set full --clock-control none --import-source on -k regex:cv_mlp_tch_kernel -s 2 -c 1 -f -o $O/m_cv_mlp_tch \
  python tools/cv_bench.py --math tch --reps 1 > $O/m_ncu_cv.log 2>&1; echo "ncu cv rc=$?"
