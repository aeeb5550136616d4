#!/bin/bash
# Round 2, call 32: evidence at HEAD: full GPU suite, smoke, bench (+ reference arm, cfg3 / cfg5 lines), ncu launch list with DRAM bytes,
# ncu --set full of the three dominant kernels, in-graph timeline, cost-volume timeline
O=gpurun_out
mkdir -p $O
rm -f $O/argmax_mismatch.log
timeout 1500 python -m pytest tests -m gpu -q -x > $O/p2_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/p2_pytest_gpu.txt
cp $O/argmax_mismatch.log $O/p2_argmax_mismatch.log 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/p2_smoke.txt 2>&1; echo "smoke rc=$?"; tail -7 $O/p2_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/p2_bench.json 2> $O/p2_bench.err; echo "bench rc=$?"; cut -c1-400 $O/p2_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/p2_bench_ref.json 2> $O/p2_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/p2_bench_ref.json
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/p2_bench_cfg3.json 2> $O/p2_bench_cfg3.err; echo "cfg3 rc=$?"; cut -c1-300 $O/p2_bench_cfg3.json
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/p2_bench_cfg5.json 2> $O/p2_bench_cfg5.err; echo "cfg5 rc=$?"; cut -c1-300 $O/p2_bench_cfg5.json
timeout 300 python tools/graph_trace.py --reps 9 --csv $O/p2_graph_trace.csv > $O/p2_graph_trace.txt 2>&1; echo "graph trace rc=$?"; head -9 $O/p2_graph_trace.txt
DTB200_DEVELOPMENT=1 DTB200_CONV_FLAGS=4096 timeout 200 python tools/cv_bench.py --math tch --reps 2 2>&1 | tail -6 | cut -c1-400 > $O/p2_cv_timeline.txt; cat $O/p2_cv_timeline.txt
timeout 200 python tools/cv_bench.py --math tch --reps 10 > $O/p2_cv_bench.txt 2>&1; cat $O/p2_cv_bench.txt
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/p2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --sustain 0 > $O/p2_bench_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l $O/p2_launches.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tch_halo_kernel -s 3 -c 1 -f -o $O/p2_conv_tch_halo python tools/conv_bench.py --math tch --only s0_64_64_3x3 --reps 3 > $O/p2_ncu1.log 2>&1; echo "ncu rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:cv_mlp_tch_kernel -s 2 -c 1 -f -o $O/p2_cv_mlp_tch python tools/cv_bench.py --math tch --reps 2 > $O/p2_ncu2.log 2>&1; echo "ncu rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tch_kernel -s 3 -c 1 -f -o $O/p2_conv_tch_s3 python tools/conv_bench.py --math tch --only s3_256_256_3x3 --reps 3 > $O/p2_ncu3.log 2>&1; echo "ncu rc=$?"
