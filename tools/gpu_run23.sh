#!/bin/bash
# 2 GPUs: split schedule + steps start when the whole call is exchanged
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_pytest.log
tail -3 gpurun_out/r2x_pytest.log
run() {  # name steps warmup
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps $2 --warmup $3 --no-eval > gpurun_out/r2x_$1_$2.json 2> gpurun_out/r2x_$1_$2.err
  python - $1 $2 <<'PY'
import json, sys
try:
    txt = open("gpurun_out/r2x_%s_%s.json" % (sys.argv[1], sys.argv[2])).read()
    j = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("RES %s steps=%s value %.1fM ms/step %.4f e2e %.1fM calls %s" % (sys.argv[1], sys.argv[2], j["value"]/1e6, j["ms_per_step"], j["e2e"]["value"]/1e6, j.get("call_ms")))
except Exception as e:
    print("RES %s ERR %s" % (sys.argv[1], e))
PY
}
export APR_BENCH_CALL_TIMES=1
for k in a b c d e f; do run all_$k 20 5; done
run all 256 32
run all 512 5
