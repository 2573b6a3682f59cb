#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; T="${1:-r02z}"
timeout 300 python tests/dev_fuse_ln.py 4 9 64 > $O/${T}_fuse_small.txt 2>&1; echo "small rc=$?"; cat $O/${T}_fuse_small.txt | tail -4
timeout 300 python tests/dev_fuse_ln.py 22 521 > $O/${T}_fuse_full.txt 2>&1; echo "full rc=$?"; cat $O/${T}_fuse_full.txt | tail -3
for f in 0 1 3; do
FP_FUSE_LN=$f timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${T}_bench_fuse$f.json 2> $O/${T}_bench_fuse$f.err; echo "bench fuse=$f rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("$O/${T}_bench_fuse$f.json") if l.startswith("{")][-1])
print("fuse=$f value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "clk", d["clocks"]["sm_mhz"])
print("   " + "  ".join("%s %.2f" % (k.replace("gemm_", ""), v["ms_per_step"]) for k, v in d["kernels"].items()))
PY
done
