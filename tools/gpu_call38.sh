#!/bin/bash
# Round 2, call 26: two MMA issuer warps in the resident halo kernel: parity, trace, in-graph A/B
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py -q -x > $O/j2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/j2_pytest.txt
timeout 100 python tools/conv_bench.py --math tch --only s0_64_64_3x3 --reps 2 --no-flush --debug 4096 2>&1 | tail -8 | cut -c1-600
timeout 300 python tools/conv_bench.py --math tch --burst 20 --no-flush --debug 0,16 --only 64_3x3 > $O/j2_conv_bench_warm.txt 2>&1; cat $O/j2_conv_bench_warm.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/j2_bench.json 2> $O/j2_bench.err; echo "bench rc=$?"
DTB200_CONV_FLAGS=16 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/j2_bench_stream.json 2> $O/j2_bench_stream.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['j2_bench','j2_bench_stream']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['value'], d['ms_per_step'], d['roofline_kernels']['conv_stack']['ms_all_launches'], d['roofline_kernels']['cost_volume_mlp_hint']['ms_per_launch'], d['sustained']['value'])
PY
