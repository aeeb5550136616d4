#!/bin/bash
# Round 2, call 29: same-box A/B of the in-graph timeline: resident (>= 3 tiles per CTA) vs streaming halo kernel; local read of the own partial
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_networks.py -q -x > $O/m2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/m2_pytest.txt
timeout 100 python tools/conv_bench.py --math tch --only s2_128_128_3x3 --reps 2 --no-flush --debug 4096 2>&1 | tail -3 | cut -c1-600
for i in 1 2; do
timeout 300 python tools/graph_trace.py --reps 9 --csv $O/m2_graph_trace_res$i.csv > $O/m2_graph_trace_res$i.txt 2>&1; head -3 $O/m2_graph_trace_res$i.txt
DTB200_CONV_FLAGS=16 timeout 300 python tools/graph_trace.py --reps 9 --csv $O/m2_graph_trace_stream$i.csv > $O/m2_graph_trace_stream$i.txt 2>&1; head -3 $O/m2_graph_trace_stream$i.txt
done
