#!/bin/bash
# Round 2, call 25: batched DSMEM reduction loads, tensormap prefetch, per-row weight barriers: parity, traces, in-graph A/B
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py tests/test_gpu_encoder.py -q -x > $O/i2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/i2_pytest.txt
for l in s0_64_64_3x3 s2_128_128_3x3 s4_384_384_3x3; do timeout 100 python tools/conv_bench.py --math tch --only $l --reps 2 --no-flush --debug 4096 2>&1 | tail -8 | cut -c1-600; done
timeout 300 python tools/conv_bench.py --math tch --burst 20 --no-flush --debug 0 > $O/i2_conv_bench_warm.txt 2>&1; cat $O/i2_conv_bench_warm.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/i2_bench.json 2> $O/i2_bench.err; echo "bench rc=$?"
DTB200_CONV_FLAGS=16 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/i2_bench_stream.json 2> $O/i2_bench_stream.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['i2_bench','i2_bench_stream']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['value'], d['ms_per_step'], d['roofline_kernels']['conv_stack']['ms_all_launches'], d['roofline_kernels']['cost_volume_mlp_hint']['ms_per_launch'], d['sustained']['value'])
PY
