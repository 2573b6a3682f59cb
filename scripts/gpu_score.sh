#!/bin/bash
# score-stage rework: parity tests, standalone timing, host profile of the video loop
set -u
O=gpurun_out; mkdir -p $O; T="${1:-r02i}"
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -x -q -k "score or fine or online or sharded or topk or video" > $O/${T}_score_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_score_tests.log
timeout 120 python tests/dev_score_bench.py > $O/${T}_score_bench.txt 2>&1; cat $O/${T}_score_bench.txt
timeout 300 python -c "
import cProfile, pstats, sys, io
sys.argv = ['bench.py', '--config', 'video', '--steps', '60', '--warmup', '3']
import bench
pr = cProfile.Profile(); pr.enable()
try:
    bench.main()
finally:
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(70); open('$O/${T}_video_hostprof.txt', 'w').write(s.getvalue())
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(40); open('$O/${T}_video_hostprof_tot.txt', 'w').write(s.getvalue())
" > $O/${T}_video_prof.json 2> $O/${T}_video_prof.err; echo "video prof rc=$?"
