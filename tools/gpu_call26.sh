#!/bin/bash
# Round 2, call 11: dot-product kernel v2 + TMA-staged variant (parity), then the north-star sweep and its ncu counters.
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_cost_volume.py tests/test_gpu_zz_simplerecon.py -q -x -k "dot or simple or behind or planes or bounds" > $O/u_pytest_dot.txt 2>&1; echo "pytest dot(ldg) rc=$?"; tail -3 $O/u_pytest_dot.txt
DTB200_CV_DOT_VARIANT=tma timeout 300 python -m pytest tests/test_gpu_cost_volume.py tests/test_gpu_zz_simplerecon.py -q -x -k "dot or simple or behind or planes or bounds" > $O/u_pytest_dot_tma.txt 2>&1; echo "pytest dot(tma) rc=$?"; tail -8 $O/u_pytest_dot_tma.txt
timeout 600 python tools/cv_sweep.py --reps 5 > $O/u_cv_sweep.txt 2>&1; echo "sweep rc=$?"; cat $O/u_cv_sweep.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:"cv_dot|cv_mlp_tch" --csv --log-file $O/u_cv_sweep_ncu.csv python tools/cv_sweep.py --once --planes 8,64 --views 1,7 > $O/u_cv_sweep_once.txt 2>&1; echo "ncu sweep rc=$?"
