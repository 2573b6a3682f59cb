#!/bin/bash
# N=8 confirmation: weak scaling (one proposal per GPU) and strong scaling (one proposal over 8 GPUs) with both exchanges
set -u
O=gpurun_out; mkdir -p $O; T="${1:-r02n}"
run() { # name, extra args...
  local name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 3 "$@" > $O/${T}_$name.json 2> $O/${T}_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/${T}_$name.json") if l.startswith("{")][-1])
    print("$name", "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "share", round(d.get("kernel_time_share_of_step", 0), 4), d.get("strong_scaling", ""))
except Exception as e:
    print("$name: no line", e)
PY
}
run weak_n8
run strong_p2p_n8 --scaling strong --exchange p2p
run strong_nccl_n8 --scaling strong --exchange nccl
