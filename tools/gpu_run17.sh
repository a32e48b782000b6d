#!/bin/bash
# the driver's SCALE command at N GPUs, both arms
N=${1:-2}
set -x
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r2q_ref_n$N.json 2> gpurun_out/r2q_ref_n$N.err ) 2> gpurun_out/r2q_ref_n$N.time; echo rc=$?
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2q_bench_n$N.json 2> gpurun_out/r2q_bench_n$N.err ) 2> gpurun_out/r2q_bench_n$N.time; echo rc=$?
cat gpurun_out/r2q_ref_n$N.time gpurun_out/r2q_bench_n$N.time; tail -c 800 gpurun_out/r2q_bench_n$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
for f in ("r2q_ref_n%s" % N, "r2q_bench_n%s" % N):
    try:
        txt = open("gpurun_out/%s.json" % f).read()
        j = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        print("SCALE", f, "value %.1fM ms/step %.4f e2e %.1fM" % (j["value"]/1e6, j["ms_per_step"], j["e2e"]["value"]/1e6), j.get("cpu_baseline", {}).get("cores"), j.get("clocks"))
        for k, v in j.get("eval", {}).items():
            print("   ", k, {kk: v[kk] for kk in ("users_per_s", "ms", "n_gpus", "k_top") if kk in v} if isinstance(v, dict) else v)
    except Exception as e:
        print("SCALE", f, "ERR", e)
PY
