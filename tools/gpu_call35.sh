#!/bin/bash
# Round 2, call 21: 256-bit epilogue accesses + unrolled issue: parity, per-layer (burst timing), trace, in-graph A/B resident vs streaming
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py -q -x > $O/e2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/e2_pytest.txt
timeout 100 python tools/conv_bench.py --math tch --only s0_64_64_3x3 --reps 2 --debug 4096 2>&1 | tail -8 | cut -c1-330
timeout 300 python tools/conv_bench.py --math tch --burst 20 --no-flush --debug 0,16 > $O/e2_conv_bench_warm.txt 2>&1; cat $O/e2_conv_bench_warm.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/e2_bench.json 2> $O/e2_bench.err; echo "bench rc=$?"
DTB200_CONV_FLAGS=16 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/e2_bench_stream.json 2> $O/e2_bench_stream.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['e2_bench','e2_bench_stream']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['value'], d['ms_per_step'], d['roofline_kernels']['conv_stack']['ms_all_launches'], d['roofline_kernels']['cost_volume_mlp_hint']['ms_per_launch'], d['sustained']['value'])
PY
