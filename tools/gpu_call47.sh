#!/bin/bash
# Round 2, call 38: cheap mbarrier poll loop (timeout check once per 4096 polls): parity, cost-volume timing, bench
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_cost_volume.py tests/test_gpu_networks.py tests/test_gpu_model.py -q -x > $O/t2_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/t2_pytest.txt
timeout 200 python tools/cv_bench.py --math tch --reps 10
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > $O/t2_bench.json 2> $O/t2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/t2_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline_kernels']['conv_stack']['ms_all_launches'], d['roofline_kernels']['cost_volume_mlp_hint']['ms_per_launch'], d['sustained']['value'])
PY
