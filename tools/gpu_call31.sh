#!/bin/bash
# Round 2, call 16: what bounds the tch halo kernel?  knock-outs (weights / patches / MMAs / stores), cold and L2-warm, + ncu --set full
O=gpurun_out
mkdir -p $O
for only in s0_64_64_3x3 s0_cat192_64_3x3 s1_64_64_3x3; do
  timeout 200 python tools/conv_bench.py --math tch --only $only --debug 0,256,512,768,1024,2048,3840 >> $O/z_knockout_cold.txt 2>&1
  timeout 200 python tools/conv_bench.py --math tch --only $only --no-flush --debug 0,256,512,768,1024,2048,3840 >> $O/z_knockout_warm.txt 2>&1
done
echo cold; cat $O/z_knockout_cold.txt; echo warm; cat $O/z_knockout_warm.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tch_halo_kernel -s 3 -c 1 -f -o $O/z_conv_tch_halo \
  python tools/conv_bench.py --math tch --only s0_64_64_3x3 --reps 3 > $O/z_ncu_conv.log 2>&1; echo "ncu rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tch_kernel -s 3 -c 1 -f -o $O/z_conv_tch_s2 \
  python tools/conv_bench.py --math tch --only s2_128_128_3x3 --reps 3 > $O/z_ncu_conv2.log 2>&1; echo "ncu rc=$?"
