#!/bin/bash
# Round 2, call 7: whole GPU suite with tch as the bench default, smoke, per-op plan profile (tch), bench with CPU baseline.
O=gpurun_out
mkdir -p $O
rm -f $O/argmax_mismatch.log
timeout 1200 python -m pytest tests -m gpu -q > $O/q_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -6 $O/q_pytest_gpu.txt
cp $O/argmax_mismatch.log $O/q_argmax_mismatch.log 2>/dev/null; grep "full" $O/q_argmax_mismatch.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/q_smoke.txt 2>&1; echo "smoke rc=$?"; tail -6 $O/q_smoke.txt
DTB200_PLAN_MATH=tch timeout 150 python tools/plan_profile.py --reps 5 --top 60 --csv $O/q_plan_ops.csv > $O/q_plan_profile.txt 2>&1; echo "plan rc=$?"; head -40 $O/q_plan_profile.txt
timeout 300 python bench.py --steps 20 --warmup 3 --cpu-budget 10 > $O/q_bench.json 2> $O/q_bench.err; echo "bench rc=$?"; cat $O/q_bench.json; tail -2 $O/q_bench.err
