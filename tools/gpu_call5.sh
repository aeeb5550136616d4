#!/bin/bash
O=gpurun_out
mkdir -p $O
# stages 4 / 3 / 2, each: full, skeleton(7936), MMA-only(3584), split-only(3328), loads-only(768)
K="0,7936,3584,3328,768,64,8000,3648,3392,832,32,7968,3616,3360,800"
for L in s0_64_64_3x3; do
  echo "#### $L"; timeout 200 python tools/conv_bench.py --only $L --debug $K 2>&1 | grep -v "^total"
done > $O/c5_stages.txt 2>&1
cat $O/c5_stages.txt
