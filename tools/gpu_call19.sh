#!/bin/bash
# Round 2, call 4: suspend-hint mbarrier waits (all tensor-core kernels), branch-free gathers in the tch volume kernel.
O=gpurun_out
mkdir -p $O
rm -f $O/argmax_mismatch.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/n_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/n_pytest_gpu.txt
cp $O/argmax_mismatch.log $O/n_argmax_mismatch.log 2>/dev/null
for m in tc3x tch; do timeout 60 python tools/cv_bench.py --math $m --reps 10; done 2>&1 | tee $O/n_cv_bench.txt
DTB200_CV_PRODUCERS=8 timeout 60 python tools/cv_bench.py --math tch --reps 10 2>&1 | tee -a $O/n_cv_bench.txt
timeout 200 python bench.py --steps 20 --warmup 3 --cpu-budget 8 > $O/n_bench.json 2> $O/n_bench.err; echo "bench rc=$?"; cut -c1-300 $O/n_bench.json; tail -2 $O/n_bench.err
timeout 100 python tools/conv_bench.py > $O/n_conv_bench.txt 2>&1; tail -16 $O/n_conv_bench.txt
timeout 120 ncu --set full --clock-control none --import-source on -k regex:cv_mlp_tch_kernel -s 2 -c 1 -f -o $O/n_cv_mlp_tch \
  python tools/cv_bench.py --math tch --reps 1 > $O/n_ncu_cv.log 2>&1; echo "ncu cv rc=$?"
