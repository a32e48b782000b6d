#!/bin/bash
# 2 GPUs: trainer exchange (symmetric-memory push vs NCCL) x schedule knobs; emulated-rank parity for the stage-0 row cache
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_parity.py -m gpu -x -q -k "shard or trainer or emul or two_gpus" > gpurun_out/r2t_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2t_pytest.log
tail -5 gpurun_out/r2t_pytest.log
run() {  # name steps warmup
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps $2 --warmup $3 --no-eval > gpurun_out/r2t_$1_$2.json 2> gpurun_out/r2t_$1_$2.err
  python - $1 $2 <<'PY'
import json, sys
try:
    txt = open("gpurun_out/r2t_%s_%s.json" % (sys.argv[1], sys.argv[2])).read()
    j = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("RES %s steps=%s value %.1fM ms/step %.4f e2e %.1fM" % (sys.argv[1], sys.argv[2], j["value"]/1e6, j["ms_per_step"], j["e2e"]["value"]/1e6))
except Exception as e:
    print("RES %s ERR %s" % (sys.argv[1], e))
PY
}
cfg() { export APR_TRAINER_EXCHANGE=$1 APR_TRAINER_ORDER=$2 APR_TRAINER_RAMP=$3 APR_TRAINER_LOOKAHEAD=$4; }
cfg symm own_first 1 1; run symm_own_r1_l1 20 5; run symm_own_r1_l1 256 32
cfg symm in_order 0 0;  run symm_ord_r0_l0 20 5; run symm_ord_r0_l0 256 32
cfg nccl in_order 0 0;  run nccl_ord_r0_l0 20 5; run nccl_ord_r0_l0 256 32
cfg symm own_first 1 0; run symm_own_r1_l0 20 5; run symm_own_r1_l0 256 32
cfg symm own_first 0 0; run symm_own_r0_l0 20 5; run symm_own_r0_l0 256 32
cfg symm own_first 1 1; run symm_own_r1_l1b 20 5
