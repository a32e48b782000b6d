#!/bin/bash
# 2 GPUs: trainer tests after the restructure, then the driver's command and a long run
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py "tests/test_gpu_parity.py::test_sharded_trainer_pipeline_single_rank" -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_pytest.log
tail -5 gpurun_out/r2r_pytest.log
for cfg in "20 5" "20 5" "256 32"; do
  set -- $cfg
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps $1 --warmup $2 --no-eval > gpurun_out/r2r_bench_n2_$1.json 2> gpurun_out/r2r_bench_n2_$1.err
  python - $1 <<'PY'
import json, sys
try:
    txt = open("gpurun_out/r2r_bench_n2_%s.json" % sys.argv[1]).read()
    j = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
    print("N2 steps=%s value %.1fM ms/step %.4f e2e %.1fM" % (sys.argv[1], j["value"]/1e6, j["ms_per_step"], j["e2e"]["value"]/1e6))
except Exception as e:
    print("N2 ERR", e)
PY
done
