#!/bin/bash
# 2 GPUs: are the cold-call stalls false dependencies between aliased streams (8 hardware connections by default)?
set -x
mkdir -p gpurun_out
run() {  # name steps warmup
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps $2 --warmup $3 --no-eval > gpurun_out/r2w_$1_$2.json 2> gpurun_out/r2w_$1_$2.err
  python - $1 $2 <<'PY'
import json, sys
try:
    txt = open("gpurun_out/r2w_%s_%s.json" % (sys.argv[1], sys.argv[2])).read()
    j = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("RES %s steps=%s value %.1fM ms/step %.4f e2e %.1fM calls %s" % (sys.argv[1], sys.argv[2], j["value"]/1e6, j["ms_per_step"], j["e2e"]["value"]/1e6, j.get("call_ms")))
except Exception as e:
    print("RES %s ERR %s" % (sys.argv[1], e))
PY
}
export APR_BENCH_CALL_TIMES=1 APR_BENCH_NO_SAMPLER=1
export CUDA_DEVICE_MAX_CONNECTIONS=32
for k in a b c d; do run conn32_split_$k 20 5; done
export APR_TRAINER_SPLIT=0 APR_TRAINER_RAMP=1 APR_TRAINER_LOOKAHEAD=1
for k in a b c; do run conn32_ramp_$k 20 5; done
unset CUDA_DEVICE_MAX_CONNECTIONS APR_TRAINER_RAMP APR_TRAINER_LOOKAHEAD
export APR_TRAINER_SPLIT=1 CUDA_MODULE_LOADING=EAGER
for k in a b c d; do run eager_split_$k 20 5; done
