#!/bin/bash
# Round 2, call 34: two issuers + 8 single-tap weight stages in the streaming halo kernel: parity, in-graph timeline, bench
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py tests/test_gpu_full_size.py -q -x > $O/q2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/q2_pytest.txt
timeout 300 python tools/graph_trace.py --reps 9 --csv $O/q2_graph_trace.csv > $O/q2_graph_trace.txt 2>&1; echo "trace rc=$?"; head -8 $O/q2_graph_trace.txt
timeout 300 python tools/conv_bench.py --math tch --burst 20 --no-flush > $O/q2_conv_bench_warm.txt 2>&1; cat $O/q2_conv_bench_warm.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/q2_bench.json 2> $O/q2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['q2_bench']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['value'], d['ms_per_step'], d['roofline_kernels']['conv_stack']['ms_all_launches'], d['roofline_kernels']['cost_volume_mlp_hint']['ms_per_launch'], d['sustained']['value'])
PY
