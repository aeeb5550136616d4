#!/bin/bash
# Round 2, call 49: capture-stream count of the DAG graph (DTB200_CONV_LANES) with the round-2 kernels: 4 / 8 (default) / 16
O=gpurun_out
mkdir -p $O
for lanes in 4 8 16; do
  DTB200_CONV_LANES=$lanes timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu --sustain 0 > $O/lanes_$lanes.json 2> $O/lanes_$lanes.err
  python - <<PY
import json
d=json.load(open('gpurun_out/lanes_$lanes.json'))
print('lanes', $lanes, d['value'], d['ms_per_step'], d['roofline_kernels']['conv_stack']['ms_all_launches'])
PY
done
