#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 600 python tools/parity_diag.py --cfg cfg5 --batch 1 --math tc3x,tch --volume-math tc3x,tch > $O/r_diag_cfg5.txt 2>&1; cat $O/r_diag_cfg5.txt | tail -24
timeout 300 python tools/parity_diag.py --cfg cfg2 --math tc3x,tch --volume-math tch > $O/r_diag_cfg2.txt 2>&1; cat $O/r_diag_cfg2.txt | tail -12
