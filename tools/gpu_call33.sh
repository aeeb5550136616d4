#!/bin/bash
# Round 2, call 18: timeline of the resident halo kernel (debug bit 0x1000), with and without knock-outs
O=gpurun_out
mkdir -p $O
for dbg in 4096 6144 5120 7936; do
  echo "== debug $dbg" >> $O/b2_trace.txt
  timeout 100 python tools/conv_bench.py --math tch --only s0_64_64_3x3 --reps 2 --debug $dbg 2>&1 | tail -24 >> $O/b2_trace.txt
done
cat $O/b2_trace.txt
