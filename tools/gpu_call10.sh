#!/bin/bash
# One gpurun call: validate HEAD on a B200 (tests, smoke, bench) + ncu launch list + two full captures.  gpurun_out/d_*.
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/d_gpu.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q > $O/d_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/d_pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/d_smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 $O/d_smoke.txt
timeout 300 python bench.py --steps 20 --warmup 3 > $O/d_bench_tc3x.json 2> $O/d_bench_tc3x.err; echo "bench rc=$?"; cut -c1-300 $O/d_bench_tc3x.json; tail -3 $O/d_bench_tc3x.err
timeout 100 python tools/conv_bench.py > $O/d_conv_bench.txt 2>&1; cat $O/d_conv_bench.txt
timeout 60 python tools/cv_bench.py --kind hint --math tc3x > $O/d_cv_bench.txt 2>&1; cat $O/d_cv_bench.txt
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
  --log-file $O/d_launches_tc3x.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/d_launches_bench.log 2>&1; echo "ncu list rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o $O/d_conv_tc_s0 \
  python tools/conv_bench.py --only s0_64_64_3x3 --reps 1 > $O/d_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:cv_mlp_tc_kernel -s 3 -c 1 -f -o $O/d_cv_mlp_tc \
  python tools/cv_bench.py --kind hint --math tc3x --reps 1 > $O/d_ncu_cv.log 2>&1; echo "ncu cv rc=$?"
ls -la $O | grep " d_"
