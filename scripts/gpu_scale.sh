#!/bin/bash
# weak + strong (peer-memory exchange) bench lines at N GPUs of one box:  bash scripts/gpu_scale.sh <tag> <N>
set -u
O=gpurun_out; mkdir -p $O; T="${1:-r03s}"; N="${2:-2}"
run() {
  local name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 20 --warmup 3 "$@" > $O/${T}_$name.json 2> $O/${T}_$name.err; echo "$name rc=$?"
  python - <<PY
import json
d = json.loads([l for l in open("$O/${T}_$name.json") if l.startswith("{")][-1])
print("$name", "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "share", round(d.get("kernel_time_share_of_step", 0), 4), "clk", d["clocks"]["sm_mhz"])
PY
}
run weak_n$N
run strong_p2p_n$N --scaling strong --exchange p2p
