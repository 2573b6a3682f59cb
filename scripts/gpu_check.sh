#!/bin/bash
# checkpoint: full GPU suite, smoke, default bench line
set -u
O=gpurun_out; mkdir -p $O; T="${1:-r02u}"
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_tests.log
timeout 300 python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${T}_smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("$O/${T}_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "clk", d["clocks"]["sm_mhz"], "frac", round(d["roofline"]["frac"], 3))
for k, v in d["kernels"].items(): print("  %-18s %7.3f ms" % (k, v["ms_per_step"]))
print(d["parity"])
PY
