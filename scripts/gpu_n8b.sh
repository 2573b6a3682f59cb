#!/bin/bash
# N=8: BASELINE configs 3-5 (ffa, video: object tracks sharded over the GPUs, refiner)
set -u
O=gpurun_out; mkdir -p $O; T="${1:-r02n8}"
run() {
  local name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 "$@" > $O/${T}_$name.json 2> $O/${T}_$name.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$O/${T}_$name.json") if l.startswith("{")][-1])
    print("$name", round(d["value"], 1), d["unit"], "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "n", d["n_gpus"])
except Exception as e:
    print("$name: no line", e); print(open("$O/${T}_$name.err").read()[-1500:])
PY
}
run video_n8 --config video --steps 300 --warmup 3
run refiner_n8 --config refiner --steps 20 --warmup 3
run ffa_n8 --config ffa --steps 25 --warmup 3
