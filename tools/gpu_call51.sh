#!/bin/bash
# Round 2, call 48: the driver's round-end sequence at the final commit: GPU suite, smoke, bench, reference arm
O=gpurun_out
mkdir -p $O
rm -f $O/argmax_mismatch.log
timeout 1500 python -m pytest tests -m gpu -q -x > $O/final_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/final_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.txt 2>&1; echo "smoke rc=$?"; tail -7 $O/final_smoke.txt
timeout 600 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"; cut -c1-200 $O/final_bench.json
