#!/bin/bash
# One gpurun call: validate HEAD on a B200 (tests, smoke, bench), per-layer timings, ncu launch list + two full captures.
# Everything lands under gpurun_out/.  If the HEAD tests fail, the same tests + bench run on the _fallback tree.
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1
RC=$?
echo "pytest rc=$RC"; tail -3 $O/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.txt
timeout 400 python bench.py --steps 20 --warmup 3 > $O/bench_tc3x.json 2> $O/bench_tc3x.err; echo "bench rc=$?"; cat $O/bench_tc3x.json | cut -c1-400
timeout 200 python tools/conv_bench.py > $O/conv_bench.txt 2>&1; cat $O/conv_bench.txt
timeout 100 python tools/cv_bench.py --kind hint --math tc3x > $O/cv_bench.txt 2>&1
timeout 100 python tools/cv_bench.py --kind dot --math exact >> $O/cv_bench.txt 2>&1
timeout 100 python tools/cv_bench.py --kind hint --math exact >> $O/cv_bench.txt 2>&1; cat $O/cv_bench.txt
if [ $RC -ne 0 ]; then
  (cd _fallback && timeout 400 python -m pytest tests -m gpu -x -q > ../$O/fb_pytest_gpu.txt 2>&1; echo "fallback pytest rc=$?"; tail -3 ../$O/fb_pytest_gpu.txt
   timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > ../$O/fb_bench_tc3x.json 2> ../$O/fb_bench.err; cut -c1-300 ../$O/fb_bench_tc3x.json
   timeout 200 python tools/conv_bench.py > ../$O/fb_conv_bench.txt 2>&1; cat ../$O/fb_conv_bench.txt)
fi
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv \
  --log-file $O/launches_tc3x.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1; echo "ncu list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 3 -c 1 -f -o $O/conv_tc_s0 \
  python tools/conv_bench.py --only s0_64_64_3x3 --reps 1 > $O/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:cv_mlp_tc_kernel -s 3 -c 1 -f -o $O/cv_mlp_tc \
  python tools/cv_bench.py --kind hint --math tc3x --reps 1 > $O/ncu_cv.log 2>&1; echo "ncu cv rc=$?"
ls -la $O
