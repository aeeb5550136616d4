#!/bin/bash
# Round 2, call 28: residual prefetch, coalesced split-K reducer, resident kernel only for >= 3 tiles per CTA: parity, timeline, bench
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py tests/test_gpu_encoder.py tests/test_gpu_full_size.py -q -x > $O/l2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/l2_pytest.txt
for l in s2_128_128_3x3 s4_384_384_3x3; do timeout 100 python tools/conv_bench.py --math tch --only $l --reps 2 --no-flush --debug 4096 2>&1 | tail -3 | cut -c1-600; done
timeout 300 python tools/graph_trace.py --csv $O/l2_graph_trace.csv > $O/l2_graph_trace.txt 2>&1; echo "trace rc=$?"; head -9 $O/l2_graph_trace.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/l2_bench.json 2> $O/l2_bench.err; echo "bench rc=$?"
DTB200_CONV_FLAGS=16 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/l2_bench_stream.json 2> $O/l2_bench_stream.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['l2_bench','l2_bench_stream']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['value'], d['ms_per_step'], d['roofline_kernels']['conv_stack']['ms_all_launches'], d['roofline_kernels']['cost_volume_mlp_hint']['ms_per_launch'], d['sustained']['value'])
PY
