#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; T="${1:-r02v}"
timeout 120 python tests/dev_ln_bench.py > $O/${T}_ln_bench.txt 2>&1; cat $O/${T}_ln_bench.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_full_config.py tests/test_refiner.py -m gpu -x -q -k "layernorm or vit or block or depth or full or forward or refiner or tokens" > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_tests.log
timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("$O/${T}_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "clk", d["clocks"]["sm_mhz"], "frac", round(d["roofline"]["frac"], 3))
for k, v in d["kernels"].items(): print("  %-18s %7.3f ms %s" % (k, v["ms_per_step"], v.get("gbs", v.get("tflops"))))
print(d["parity"])
PY
