#!/bin/bash
# Round 2, call 15: cluster split-K (DSMEM reduction) + programmatic dependent launch in the tch conv kernels: parity, then A/B bench.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py tests/test_gpu_encoder.py tests/test_gpu_full_size.py -q -x > $O/y_pytest.txt 2>&1; echo "pytest rc=$?"; tail -15 $O/y_pytest.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/y_bench_pdl.json 2> $O/y_bench_pdl.err; echo "bench rc=$?"; cat $O/y_bench_pdl.json
DTB200_CONV_PDL=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/y_bench_nopdl.json 2> $O/y_bench_nopdl.err; echo "bench(no pdl) rc=$?"; cat $O/y_bench_nopdl.json
timeout 300 python tools/plan_profile.py --csv $O/y_plan_ops.csv > $O/y_plan_profile.txt 2>&1; echo "plan profile rc=$?"; head -8 $O/y_plan_profile.txt
