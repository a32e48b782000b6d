#!/bin/bash
# round-2 GPU session 7 (8 GPUs): resident-block / launch-order sweep of the sharded step (train only, short runs)
N=${1:-8}
set -x
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 64 --warmup 16 --no-eval > gpurun_out/r2f_n${N}_$tag.json 2> gpurun_out/r2f_n${N}_$tag.err
  python - gpurun_out/r2f_n${N}_$tag.json $tag <<'PY'
import json, sys
try:
    txt = open(sys.argv[1]).read()
    j = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("SWEEP", sys.argv[2], "value %.0fM ms/step %.4f e2e %.0fM" % (j["value"]/1e6, j["ms_per_step"], j["e2e"]["value"]/1e6))
except Exception as e:
    print("SWEEP", sys.argv[2], "ERR", e)
PY
}
run base APR_DUMMY=0
run gen4 APR_GEN_BLOCKS=4
run gen4_fast1 APR_GEN_BLOCKS=4 APR_FAST_BLOCKS=1
run gen8_fast1 APR_GEN_BLOCKS=8 APR_FAST_BLOCKS=1
run order1_gen4 APR_SHARD_ORDER=1 APR_GEN_BLOCKS=4
run nopairs_gen4 APR_PAIRS=0 APR_GEN_BLOCKS=4
