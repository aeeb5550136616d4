#!/bin/bash
# Round 2, call 44: final bench lines at HEAD (default, reference arm)
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 3 > $O/w2_bench.json 2> $O/w2_bench.err; echo "bench rc=$?"; cut -c1-300 $O/w2_bench.json; tail -3 $O/w2_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/w2_bench_ref.json 2> $O/w2_bench_ref.err; echo "ref rc=$?"; cut -c1-200 $O/w2_bench_ref.json
