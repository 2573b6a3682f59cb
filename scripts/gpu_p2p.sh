#!/bin/bash
# N=2 check of the peer-memory score exchange: the two strong-scaling tests, then the full-size strong bench with both exchanges.
set -u
O=gpurun_out; mkdir -p $O; T="${1:-r02h}"
timeout 600 python -m pytest tests/test_bench_contract.py -m gpu -x -q -k strong > $O/${T}_p2p_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${T}_p2p_tests.log
for ex in p2p nccl; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --scaling strong --exchange $ex --steps 20 --warmup 3 > $O/${T}_strong_n2_$ex.json 2> $O/${T}_strong_n2_$ex.err; echo "strong $ex rc=$?"
  tail -c 600 $O/${T}_strong_n2_$ex.json
done
