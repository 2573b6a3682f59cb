#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; T="${1:-r02j}"
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -x -q -k "score or fine or online or sharded or retriev or topk" > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_tests.log
timeout 120 python tests/dev_score_bench.py > $O/${T}_score_bench.txt 2>&1; cat $O/${T}_score_bench.txt
timeout 200 python tests/dev_retrieval_bench.py > $O/${T}_retrieval_bench.txt 2>&1; head -6 $O/${T}_retrieval_bench.txt
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k "regex:score_rows" -c 2 python tests/dev_score_bench.py 2>&1 | grep -E "score_rows|gpu__time|dram__bytes|issue_active|inst_executed" | head -20 > $O/${T}_score_ncu.txt; cat $O/${T}_score_ncu.txt
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k "regex:scan_kernel" -s 30 -c 60 python tests/dev_retrieval_bench.py 2>&1 | grep -E "gpu__time" | awk '{print $3}' | tr '\n' ' ' > $O/${T}_scan_ncu.txt; cat $O/${T}_scan_ncu.txt
