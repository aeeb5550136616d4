#!/bin/bash
# Round 2, call 35: cfg4 (ScanNetv2 test split shapes, sharded keyframes) at N = 1; DRAM traffic of the step with caches kept between kernels
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py --workload cfg4 --steps 64 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/r2_bench_cfg4_n1.json 2> $O/r2_bench_cfg4_n1.err; echo "cfg4 rc=$?"; cut -c1-700 $O/r2_bench_cfg4_n1.json; tail -3 $O/r2_bench_cfg4_n1.err
timeout 900 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2_launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --sustain 0 > $O/r2_bench_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l $O/r2_launches_warm.csv
