#!/bin/bash
# round-2 GPU session 4 (2 GPUs): real 2-rank tests, pipelined sharded trainer, bench at N=2 (train + item-sharded eval)
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2d_smi.txt
timeout 900 python -m pytest tests/test_gpu_distributed.py "tests/test_gpu_parity.py::test_sharded_trainer_pipeline_single_rank" -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_pytest.log
tail -25 gpurun_out/r2d_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 256 --warmup 32 --eval-users 131072 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; echo rc=$?
tail -c 2000 gpurun_out/r2d_bench_n2.err
APR_SHARD_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 64 --warmup 16 --no-eval > gpurun_out/r2d_bench_n2_timing.json 2> gpurun_out/r2d_bench_n2_timing.err; echo rc=$?
grep "shard timing" gpurun_out/r2d_bench_n2_timing.err | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2d_ref_n2.json 2> gpurun_out/r2d_ref_n2.err; echo rc=$?
python - <<'PY'
import json
for f in ("r2d_bench_n2", "r2d_bench_n2_timing", "r2d_ref_n2"):
    try:
        j = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.0fM ms/step %.4f e2e %.0fM" % (j["value"]/1e6, j["ms_per_step"], j["e2e"]["value"]/1e6), j.get("roofline", {}).get("nvlink_gb_s_per_direction_per_gpu"), j.get("cpu_baseline", {}).get("cores"))
        if "eval" in j: print(json.dumps(j["eval"], indent=1)[:2500])
    except Exception as e:
        print(f, "ERR", e)
PY
