#!/bin/bash
# Round 2, call 27: in-graph timeline of the cfg-2 conv plan, resident vs streaming halo kernel
O=gpurun_out
mkdir -p $O
timeout 300 python tools/graph_trace.py --csv $O/k2_graph_trace.csv > $O/k2_graph_trace.txt 2>&1; echo "trace rc=$?"; head -14 $O/k2_graph_trace.txt
DTB200_CONV_FLAGS=16 timeout 300 python tools/graph_trace.py --csv $O/k2_graph_trace_stream.csv > $O/k2_graph_trace_stream.txt 2>&1; echo "trace rc=$?"; head -14 $O/k2_graph_trace_stream.txt
DTB200_CONV_PDL=0 timeout 300 python tools/graph_trace.py --csv $O/k2_graph_trace_nopdl.csv > $O/k2_graph_trace_nopdl.txt 2>&1; echo "trace rc=$?"; head -14 $O/k2_graph_trace_nopdl.txt
