#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; T="${1:-r02o}"
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_refiner.py tests/test_template_store.py -m gpu -x -q -k "raster or render or point or clip or near or refiner or pipeline or smoke or template or forward" > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_tests.log
for big in 8 16 32 64; do echo "big=$big"; FP_RASTER_BIG=$big timeout 200 python tests/dev_raster_perf.py 2>&1 | tail -1; done > $O/${T}_raster_perf.txt; cat $O/${T}_raster_perf.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:clear_keys|triangle_kernel|resolve_kernel|vertex_kernel|crop_kernel|mask_bbox" -c 14 python tests/dev_raster_once.py 2>&1 | grep -E "^\s+(void )?(fp::|unnamed)|gpu__time|dram__bytes" > $O/${T}_raster_ncu.txt; grep -E "kernel|gpu__time" $O/${T}_raster_ncu.txt | cut -c1-90 | tail -10
