#!/bin/bash
# Round 2, call 2: first run of the persistent fp16-split feature-volume kernel (math = tch): parity, then timing vs tc3x.
O=gpurun_out
mkdir -p $O
rm -f $O/argmax_mismatch.log
timeout 300 python -m pytest tests/test_gpu_cost_volume.py -q -x -k "tch" > $O/l_pytest_tch.txt 2>&1; echo "pytest tch(16) rc=$?"; tail -15 $O/l_pytest_tch.txt
DTB200_CV_PRODUCERS=8 timeout 300 python -m pytest tests/test_gpu_cost_volume.py -q -x -k "tch" > $O/l_pytest_tch8.txt 2>&1; echo "pytest tch(8) rc=$?"; tail -5 $O/l_pytest_tch8.txt
cp $O/argmax_mismatch.log $O/l_argmax_mismatch.log 2>/dev/null
for m in tc3x tch; do timeout 60 python tools/cv_bench.py --math $m --reps 10; done 2>&1 | tee $O/l_cv_bench.txt
DTB200_CV_PRODUCERS=8 timeout 60 python tools/cv_bench.py --math tch --reps 10 2>&1 | tee -a $O/l_cv_bench.txt
timeout 200 python -m pytest tests/test_gpu_full_size.py -q -x -k "cfg2" > $O/l_pytest_full.txt 2>&1; echo "pytest full rc=$?"; tail -5 $O/l_pytest_full.txt
timeout 100 python -m pytest tests/test_gpu_tsdf.py -q -k "torch_cuda" > $O/l_pytest_tsdf.txt 2>&1; echo "pytest tsdf rc=$?"; grep "tsdf aten_cuda" $O/l_pytest_tsdf.txt | grep -v print; tail -3 $O/l_pytest_tsdf.txt
