#!/bin/bash
# 2-GPU call: torchrun bench at N=2 (sharded frames + NCCL depth-map gather), reference arm rank handling, new full-size tests.
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 3 > $O/e_bench_n2.json 2> $O/e_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-400 $O/e_bench_n2.json; tail -5 $O/e_bench_n2.err
timeout 300 python -m pytest tests/test_gpu_full_size.py -x -q > $O/e_pytest_full.txt 2>&1; echo "full-size rc=$?"; tail -15 $O/e_pytest_full.txt
