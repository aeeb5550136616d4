#!/bin/bash
# Round 2, call 17: resident-weights halo kernel: parity, per-layer timing with knock-outs, bench
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py tests/test_gpu_full_size.py -q -x > $O/a2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/a2_pytest.txt
for only in s0_64_64_3x3 s0_24_64 s1_64_64_3x3; do
  timeout 200 python tools/conv_bench.py --math tch --only $only --debug 0,16,256,512,1024,2048,3840 >> $O/a2_knockout_cold.txt 2>&1
done
cat $O/a2_knockout_cold.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/a2_bench.json 2> $O/a2_bench.err; echo "bench rc=$?"; cat $O/a2_bench.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tch_halo_res_kernel -s 3 -c 1 -f -o $O/a2_conv_tch_halo_res \
  python tools/conv_bench.py --math tch --only s0_64_64_3x3 --reps 3 > $O/a2_ncu_conv.log 2>&1; echo "ncu rc=$?"
