#!/bin/bash
# Round 2, call 1: fp16-split probe, the never-run tail-split kernel, full GPU suite with arg-max mismatch counts and the
# full-size cfg3/cfg4/cfg5 cases, smoke, per-op plan profile, TSDF device timing, HEAD bench, ncu of the default halo kernel.
O=gpurun_out
mkdir -p $O
rm -f $O/argmax_mismatch.log
timeout 60 tools/f16_probe.bin > $O/k_f16_probe.txt 2>&1; echo "f16 probe rc=$?"; cat $O/k_f16_probe.txt
DTB200_TEST_UNVALIDATED=1 DTB200_CONV_WS_SLOTS=24 timeout 120 python -m pytest tests/test_gpu_zz_halo_tail.py -q -x > $O/k_pytest_tail.txt 2>&1; echo "tail rc=$?"; tail -5 $O/k_pytest_tail.txt
timeout 900 python -m pytest tests -m gpu -q > $O/k_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -8 $O/k_pytest_gpu.txt
cp $O/argmax_mismatch.log $O/k_argmax_mismatch.log 2>/dev/null; cat $O/k_argmax_mismatch.log | head -60
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/k_smoke.txt 2>&1; echo "smoke rc=$?"; tail -5 $O/k_smoke.txt
timeout 150 python tools/plan_profile.py --reps 5 --top 200 --csv $O/k_plan_ops.csv > $O/k_plan_profile.txt 2>&1; echo "plan rc=$?"; tail -25 $O/k_plan_profile.txt
timeout 60 python tools/tsdf_bench.py --reps 5 > $O/k_tsdf_bench.txt 2>&1; echo "tsdf rc=$?"; tail -8 $O/k_tsdf_bench.txt
timeout 200 python bench.py --steps 20 --warmup 3 --cpu-budget 10 > $O/k_bench.json 2> $O/k_bench.err; echo "bench rc=$?"; cut -c1-300 $O/k_bench.json
DTB200_CONV_FLAGS=16 DTB200_CONV_WS_SLOTS=24 timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/k_bench_tail.json 2> $O/k_bench_tail.err; echo "bench tail rc=$?"; cut -c1-200 $O/k_bench_tail.json
timeout 100 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo_kernel -s 3 -c 1 -f -o $O/k_conv_tc_halo3 \
  python tools/conv_bench.py --only s0_64_64_3x3 --reps 1 > $O/k_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
