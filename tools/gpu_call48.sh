#!/bin/bash
# Round 2, call 41: cluster split-K through an L2 scratch, staged rows copied out coalesced, batched reduce loads: parity, timeline, A/B bench
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py tests/test_gpu_encoder.py tests/test_gpu_full_size.py -q -x > $O/u2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/u2_pytest.txt
for l in s2_128_128_3x3 s4_384_384_3x3; do timeout 100 python tools/conv_bench.py --math tch --only $l --reps 2 --no-flush --debug 4096 2>&1 | tail -3 | cut -c1-600; done
timeout 300 python tools/conv_bench.py --math tch --burst 20 --no-flush --debug 0,32 --only s > $O/u2_conv_bench_warm.txt 2>&1; cat $O/u2_conv_bench_warm.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/u2_bench.json 2> $O/u2_bench.err; echo "bench rc=$?"
DTB200_CONV_FLAGS=32 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/u2_bench_dsmem.json 2> $O/u2_bench_dsmem.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['u2_bench','u2_bench_dsmem']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['value'], d['ms_per_step'], d['roofline_kernels']['conv_stack']['ms_all_launches'], d['roofline_kernels']['cost_volume_mlp_hint']['ms_per_launch'], d['sustained']['value'])
PY
