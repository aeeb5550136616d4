#!/bin/bash
# Round 2, call 12: TSDF ray-cast hint + incremental loop; north-star sweep with the TMA-staged dot variant (rebuilt library).
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_tsdf.py -q -x > $O/v_pytest_tsdf.txt 2>&1; echo "pytest tsdf rc=$?"; tail -12 $O/v_pytest_tsdf.txt
DTB200_CV_DOT_VARIANT=tma timeout 300 python -m pytest tests/test_gpu_cost_volume.py tests/test_gpu_zz_simplerecon.py -q -x -k "dot or simple or behind or planes or bounds" > $O/v_pytest_dot_tma.txt 2>&1; echo "pytest dot(tma) rc=$?"; tail -3 $O/v_pytest_dot_tma.txt
timeout 900 python tools/cv_sweep.py --reps 5 > $O/v_cv_sweep.txt 2>&1; echo "sweep rc=$?"; cat $O/v_cv_sweep.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:"cv_dot|cv_mlp_tch" --csv --log-file $O/v_cv_sweep_ncu.csv python tools/cv_sweep.py --once --planes 8,64 --views 1,7 > $O/v_cv_sweep_once.txt 2>&1; echo "ncu sweep rc=$?"
timeout 60 python tools/tsdf_bench.py --reps 5 > $O/v_tsdf_bench.txt 2>&1; echo "tsdf bench rc=$?"; tail -6 $O/v_tsdf_bench.txt
