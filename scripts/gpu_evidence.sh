#!/bin/bash
# Round-2 evidence run (under gpurun, one GPU): tests, bench lines, ncu launch list + full captures, sanitizer.
# Outputs land in gpurun_out/; the summaries are produced here and copied into profiles/ afterwards.
set -u
O=gpurun_out
mkdir -p $O
T="${1:-r02zz}"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 700 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -2 $O/${T}_tests.log
fi
# (.ncu-rep files are summarised on the box and deleted: gpurun only brings back 64 MiB)
summ() { python profiles/summarize.py full $O/$1.ncu-rep $O/$1_full.txt && rm -f $O/$1.ncu-rep; }
timeout 300 python bench.py --steps 20 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --res 420 --hyp 600 --chunk 150 --no-cpu-baseline > $O/${T}_bench_420.json 2>/dev/null; echo "bench420 rc=$?"
timeout 300 python bench.py --config ffa --steps 196 --warmup 3 > $O/${T}_ffa.json 2>/dev/null; echo "ffa rc=$?"
timeout 300 python bench.py --config video --steps 300 --warmup 3 > $O/${T}_video.json 2>/dev/null; echo "video rc=$?"
timeout 300 python bench.py --config refiner --steps 20 --warmup 3 > $O/${T}_refiner.json 2>/dev/null; echo "refiner rc=$?"
# ncu: launch list of one full-depth step (shares), then full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_launches.csv python profiles/prof_step.py 2 520 22 > /dev/null 2>&1; echo "launches rc=$?"
timeout 900 ncu --set full --clock-control none -o $O/${T}_step1layer -f python profiles/prof_step.py 1 520 1 > $O/${T}_ncu_step.log 2>&1; echo "ncu step rc=$?"; summ ${T}_step1layer
timeout 300 ncu --set full --clock-control none -k "regex:score_rows|layernorm_kernel" -s 4 -c 2 -o $O/${T}_score_ln -f python tests/dev_score_ln_once.py > $O/${T}_ncu_score.log 2>&1; echo "ncu score/ln rc=$?"; summ ${T}_score_ln
ATTN_B=150 timeout 300 ncu --set full --clock-control none -k regex:attention_pair -s 2 -c 1 -o $O/${T}_attention_pair -f python tests/dev_attn_bench.py 905 > $O/${T}_ncu_pair.log 2>&1; echo "ncu pair rc=$?"; summ ${T}_attention_pair
timeout 300 ncu --set full --clock-control none -k "regex:scan_kernel|topk_large|fine_kernel" -c 3 -o $O/${T}_retrieval -f python tests/dev_retrieval_bench.py > $O/${T}_ncu_retr.log 2>&1; echo "ncu retrieval rc=$?"; summ ${T}_retrieval
FP_RASTER_TILE=1 timeout 300 ncu --set full --clock-control none -k "regex:tile_kernel|bin_kernel|tile_scan" -s 4 -c 4 -o $O/${T}_raster_tile -f python tests/dev_raster_once.py > $O/${T}_ncu_tile.log 2>&1; echo "ncu tile rc=$?"; summ ${T}_raster_tile
# compute-sanitizer over every kernel family (tests/dev_sanitize.py)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tests/dev_sanitize.py > $O/${T}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 $O/${T}_memcheck.log
# score stage and rasteriser: kernel-level times and DRAM bytes
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k "regex:score_rows|score_reduce|prep_query|topk_kernel" -c 8 python tests/dev_score_bench.py 2>&1 | grep -E "^\s+(void )?(fp::|unnamed)|gpu__time|dram__bytes|issue_active" > $O/${T}_score_ncu.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:clear_keys|triangle_kernel|resolve_kernel|vertex_kernel|crop_kernel|mask_bbox" -c 14 python tests/dev_raster_once.py 2>&1 | grep -E "^\s+(void )?(fp::|unnamed)|gpu__time|dram__bytes" > $O/${T}_raster_ncu.txt
python profiles/summarize.py launches $O/${T}_launches.csv $O/${T}_launches_summary.txt; rm -f $O/${T}_launches.csv
rm -f $O/*.ncu-rep; du -sh $O
