#!/bin/bash
# 2 GPUs: where does the time of a short timed region go?  per-call times, with and without the nvidia-smi sampler
set -x
mkdir -p gpurun_out
run() {  # name steps warmup
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps $2 --warmup $3 --no-eval > gpurun_out/r2u_$1_$2.json 2> gpurun_out/r2u_$1_$2.err
  python - $1 $2 <<'PY'
import json, sys
try:
    txt = open("gpurun_out/r2u_%s_%s.json" % (sys.argv[1], sys.argv[2])).read()
    j = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("RES %s steps=%s value %.1fM ms/step %.4f e2e %.1fM calls %s" % (sys.argv[1], sys.argv[2], j["value"]/1e6, j["ms_per_step"], j["e2e"]["value"]/1e6, j.get("call_ms")))
except Exception as e:
    print("RES %s ERR %s" % (sys.argv[1], e))
PY
}
export APR_BENCH_CALL_TIMES=1
export APR_TRAINER_EXCHANGE=symm APR_TRAINER_ORDER=own_first APR_TRAINER_RAMP=0 APR_TRAINER_LOOKAHEAD=0
APR_BENCH_NO_SAMPLER=0 run samp_r0 512 5
APR_BENCH_NO_SAMPLER=1 run nosamp_r0 512 5
APR_BENCH_NO_SAMPLER=1 run nosamp_r0 20 5
APR_BENCH_NO_SAMPLER=1 run nosamp_r0b 20 5
APR_BENCH_NO_SAMPLER=0 run samp_r0 20 5
export APR_TRAINER_RAMP=1 APR_TRAINER_LOOKAHEAD=1
APR_BENCH_NO_SAMPLER=1 run nosamp_r1 512 5
APR_BENCH_NO_SAMPLER=1 run nosamp_r1 20 5
APR_BENCH_NO_SAMPLER=0 run samp_r1 20 5
