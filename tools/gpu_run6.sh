#!/bin/bash
# round-2 GPU session 6 (N GPUs, default 8): bench at N (train + item-sharded eval) and the per-stage timing table
N=${1:-8}
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_smi_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 128 --warmup 16 --eval-users 262144 > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err; echo rc=$?
tail -c 1500 gpurun_out/r2e_bench_n$N.err
APR_SHARD_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 32 --warmup 16 --no-eval > gpurun_out/r2e_bench_n${N}_timing.json 2> gpurun_out/r2e_bench_n${N}_timing.err; echo rc=$?
grep "shard timing" gpurun_out/r2e_bench_n${N}_timing.err | tail -4
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for f in ("r2e_bench_n%s" % N, "r2e_bench_n%s_timing" % N):
    try:
        txt = open("gpurun_out/%s.json" % f).read()
        j = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print(f, "value %.0fM ms/step %.4f e2e %.0fM nvlink %.0f GB/s/dir" % (j["value"]/1e6, j["ms_per_step"], j["e2e"]["value"]/1e6, j["roofline"]["nvlink_gb_s_per_direction_per_gpu"]), j["config"].get("steps_per_call"))
        for k, v in j.get("eval", {}).items():
            print(k, {kk: v[kk] for kk in ("users_per_s", "ms", "n_gpus", "k_top", "whole_run_frac_of_sustained_peak") if kk in v} if isinstance(v, dict) else v)
    except Exception as e:
        print(f, "ERR", e)
PY
