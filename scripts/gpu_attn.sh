#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; T="${1:-r02q}"
for d in 0 800 1500 2200; do echo "delay1=$d"; FP_ATTN_DELAY1=$d timeout 120 python tests/dev_attn_phases.py 2>&1 | tail -7; done > $O/${T}_attn_phases.txt 2>&1; cat $O/${T}_attn_phases.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "attention" > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_tests.log
