#!/bin/bash
# HEAD validation with the halo-tile kernel on by default + A/B of the 3-taps-per-weight-stage variant.  gpurun_out/h_*.
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/h_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/h_pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/h_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/h_smoke.txt
timeout 200 python bench.py --steps 20 --warmup 3 > $O/h_bench_tc3x.json 2> $O/h_bench_tc3x.err; echo "bench rc=$?"; cut -c1-200 $O/h_bench_tc3x.json; tail -3 $O/h_bench_tc3x.err
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv \
  --log-file $O/h_launches_tc3x.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/h_launches_bench.log 2>&1; echo "ncu list rc=$?"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:conv_tc_halo_kernel -s 3 -c 1 -f -o $O/h_conv_tc_halo \
  python tools/conv_bench.py --only s0_64_64_3x3 --reps 1 > $O/h_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 100 python tools/conv_bench.py --debug 0,8,128 > $O/h_conv_bench.txt 2>&1; cat $O/h_conv_bench.txt
DTB200_CONV_FLAGS=8 timeout 200 python -m pytest tests/test_gpu_networks.py tests/test_gpu_full_size.py -x -q > $O/h_pytest_t3.txt 2>&1; echo "pytest t3 rc=$?"; tail -3 $O/h_pytest_t3.txt
DTB200_CONV_FLAGS=8 timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/h_bench_t3.json 2> $O/h_bench_t3.err; echo "bench t3 rc=$?"; cut -c1-200 $O/h_bench_t3.json
