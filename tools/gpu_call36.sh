#!/bin/bash
# Round 2, call 23: leaner conv epilogue (activation as template parameter, saturating pack): parity, trace, in-graph A/B
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_networks.py tests/test_gpu_model.py -q -x > $O/g2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/g2_pytest.txt
for d in 4096 12288; do timeout 100 python tools/conv_bench.py --math tch --only s0_64_64_3x3 --reps 2 --debug $d 2>&1 | tail -8 | cut -c1-330; done
timeout 300 python tools/conv_bench.py --math tch --burst 20 --no-flush --debug 0,16 > $O/g2_conv_bench_warm.txt 2>&1; cat $O/g2_conv_bench_warm.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/g2_bench.json 2> $O/g2_bench.err; echo "bench rc=$?"
DTB200_CONV_FLAGS=16 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/g2_bench_stream.json 2> $O/g2_bench_stream.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['g2_bench','g2_bench_stream']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['value'], d['ms_per_step'], d['roofline_kernels']['conv_stack']['ms_all_launches'], d['roofline_kernels']['cost_volume_mlp_hint']['ms_per_launch'], d['sustained']['value'])
PY
