#!/bin/bash
# Round 2, call 31: cost-volume tch kernel: split hint MLP over both epilogue halves (16-byte weight loads, prefetched hint inputs),
# per-pixel producer state reused across plane chunks: parity, timeline, timing, bench
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_cost_volume.py tests/test_gpu_model.py tests/test_gpu_full_size.py -q -x > $O/o2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/o2_pytest.txt
DTB200_DEVELOPMENT=1 DTB200_CONV_FLAGS=4096 timeout 200 python tools/cv_bench.py --math tch --reps 2 2>&1 | tail -6 | cut -c1-400
timeout 200 python tools/cv_bench.py --math tch --reps 10
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/o2_bench.json 2> $O/o2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ['o2_bench']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['value'], d['ms_per_step'], d['roofline_kernels']['conv_stack']['ms_all_launches'], d['roofline_kernels']['cost_volume_mlp_hint']['ms_per_launch'], d['sustained']['value'])
PY
