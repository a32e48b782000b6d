#!/bin/bash
# round-2 GPU session 14 (8 GPUs): what the remote workspace REDs cost (APR_STEP_FLAGS=1: timing only, wrong results)
N=${1:-8}
set -x
mkdir -p gpurun_out
for fl in 1 0; do
APR_STEP_FLAGS=$fl APR_SHARD_TIMING=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$fl bench.py --gpus $N --steps 32 --warmup 16 --no-eval > gpurun_out/r2m_n${N}_flags$fl.json 2> gpurun_out/r2m_n${N}_flags$fl.err
echo "FLAGS=$fl"; grep "shard timing" gpurun_out/r2m_n${N}_flags$fl.err | tail -2
python - gpurun_out/r2m_n${N}_flags$fl.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    print("RESULT", sys.argv[1], "value %.0fM ms/step %.4f" % (j["value"]/1e6, j["ms_per_step"]))
except Exception as e:
    print("RESULT ERR", e)
PY
done
