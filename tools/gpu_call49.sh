#!/bin/bash
# Round 2, call 43: HEAD after the head-kernel rewrite and the bench changes (burst-peak fraction, cfg4): full GPU suite, smoke, bench, timeline
O=gpurun_out
mkdir -p $O
rm -f $O/argmax_mismatch.log
timeout 1500 python -m pytest tests -m gpu -q -x > $O/v2_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/v2_pytest_gpu.txt
cp $O/argmax_mismatch.log $O/v2_argmax_mismatch.log 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/v2_smoke.txt 2>&1; echo "smoke rc=$?"; tail -7 $O/v2_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 3 > $O/v2_bench.json 2> $O/v2_bench.err; echo "bench rc=$?"; cut -c1-300 $O/v2_bench.json
timeout 300 python tools/graph_trace.py --reps 9 --csv $O/v2_graph_trace.csv > $O/v2_graph_trace.txt 2>&1; echo "graph trace rc=$?"; head -8 $O/v2_graph_trace.txt; tail -3 $O/v2_graph_trace.txt
