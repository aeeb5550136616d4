#!/bin/bash
# Round 2, call 9: ring-hazard fix in the tch volume kernel (per-writer-group empty barriers); new bench fields.
O=gpurun_out
mkdir -p $O
rm -f $O/argmax_mismatch.log
timeout 600 python -m pytest tests/test_gpu_cost_volume.py -q -x > $O/s_pytest_cv.txt 2>&1; echo "pytest cv rc=$?"; tail -4 $O/s_pytest_cv.txt
timeout 600 python tools/parity_diag.py --cfg cfg5 --batch 1 --math tch --volume-math tch > $O/s_diag_cfg5.txt 2>&1; tail -5 $O/s_diag_cfg5.txt | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_full_size.py -q > $O/s_pytest_full.txt 2>&1; echo "pytest full rc=$?"; tail -4 $O/s_pytest_full.txt
timeout 60 python tools/cv_bench.py --math tch --reps 10 2>&1 | tee $O/s_cv_bench.txt
timeout 400 python bench.py --steps 20 --warmup 3 --cpu-budget 12 > $O/s_bench.json 2> $O/s_bench.err; echo "bench rc=$?"; cat $O/s_bench.json; tail -2 $O/s_bench.err
