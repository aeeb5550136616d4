#!/bin/bash
# Round 2, call 14: HEAD evidence: full GPU suite, smoke, default bench (+cpu baseline, sustained, reference-on-GPU), reference arm,
# ncu launch list with DRAM bytes (uncut), per-op plan profile.
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/x_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/x_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/x_smoke.txt 2>&1; echo "smoke rc=$?"; tail -6 $O/x_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/x_bench.json 2> $O/x_bench.err; echo "bench rc=$?"; cat $O/x_bench.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/x_bench_ref.json 2> $O/x_bench_ref.err; echo "ref rc=$?"; cat $O/x_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/x_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu --sustain 0 > $O/x_bench_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l $O/x_launches.csv
timeout 300 python tools/plan_profile.py > $O/x_plan_profile.txt 2>&1; echo "plan profile rc=$?"; tail -40 $O/x_plan_profile.txt
