#!/bin/bash
for d in 0 4 12; do FP_GEMM_DEBUG=$d timeout 200 python tests/dev_fuse_ln.py 22 521 2>&1 | tail -1; done
