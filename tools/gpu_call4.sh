#!/bin/bash
O=gpurun_out
mkdir -p $O
K="0,256,512,768,1024,2048,3072,3328,3584,4096,7936"
for L in s0_64_64_3x3 s0_cat192_64_3x3 s2_128_128_3x3 s1_64_64_3x3; do
  echo "#### $L"; timeout 200 python tools/conv_bench.py --only $L --debug $K 2>&1 | grep -v "^total"
done > $O/c4_knockout.txt 2>&1
cat $O/c4_knockout.txt
