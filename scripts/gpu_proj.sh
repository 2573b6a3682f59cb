#!/bin/bash
for d in 0 1 2 3; do echo "FP_GEMM_PREFETCH=$d"; FP_GEMM_PREFETCH=$d timeout 100 python tests/dev_proj_probe.py 2>&1 | tail -3; done
