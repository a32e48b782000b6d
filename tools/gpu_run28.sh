#!/bin/bash
# 4 GPUs: the trainer's symmetric-memory exchange at G = 4, driver command, and with the global batch of the 8-GPU run
set -x
mkdir -p gpurun_out
run() {  # name steps warmup extra
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps $2 --warmup $3 --no-eval $4 > gpurun_out/r2ac_$1_$2.json 2> gpurun_out/r2ac_$1_$2.err
  python - $1 $2 <<'PY'
import json, sys
try:
    txt = open("gpurun_out/r2ac_%s_%s.json" % (sys.argv[1], sys.argv[2])).read()
    j = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("RES %s steps=%s value %.1fM ms/step %.4f e2e %.1fM calls %s" % (sys.argv[1], sys.argv[2], j["value"]/1e6, j["ms_per_step"], j["e2e"]["value"]/1e6, j.get("call_ms")))
except Exception as e:
    print("RES %s ERR %s" % (sys.argv[1], e))
    print(open("gpurun_out/r2ac_%s_%s.err" % (sys.argv[1], sys.argv[2])).read()[-1500:])
PY
}
export APR_BENCH_CALL_TIMES=1
run n4 20 5
run n4 256 32
run n4_b131072 20 5 "--batch 131072"
