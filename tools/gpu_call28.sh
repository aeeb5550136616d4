#!/bin/bash
# Round 2, call 13: matching encoder on device (row N1): parity vs reference fixtures + end to end from images.
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_encoder.py -q -x > $O/w_pytest_enc.txt 2>&1; echo "pytest encoder rc=$?"; tail -30 $O/w_pytest_enc.txt
