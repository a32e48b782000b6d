#!/bin/bash
# round-2 GPU session 2: tensor-core top-k / merge / sharded-table tests, graph-cache bench check, eval bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval_tc.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -30 gpurun_out/r2b_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-eval --no-variants --no-cpu > gpurun_out/r2b_bench_20_5.json 2> gpurun_out/r2b_bench_20_5.err; echo rc=$?
timeout 600 python bench.py --gpus 1 --steps 2048 --warmup 128 --no-eval --no-variants --no-cpu > gpurun_out/r2b_bench_2048.json 2> gpurun_out/r2b_bench_2048.err; echo rc=$?
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-variants --no-cpu --eval-users 131072 > gpurun_out/r2b_bench_eval.json 2> gpurun_out/r2b_bench_eval.err; echo rc=$?
tail -c 1500 gpurun_out/r2b_bench_eval.err
python - <<'PY'
import json
for f in ("r2b_bench_20_5", "r2b_bench_2048"):
    try:
        j = json.load(open("gpurun_out/%s.json" % f)); r = j["roofline"]
        print(f, "value %.0fM ms %.4f kern %.4f frac %.3f whole %.3f e2e %.0fM" % (j["value"]/1e6, j["ms_per_step"], r["ms_per_step_kernel"], r["frac"], r["whole_step_frac"], j["e2e"]["value"]/1e6), j["graph_cache"])
    except Exception as e:
        print(f, "ERR", e)
try:
    j = json.load(open("gpurun_out/r2b_bench_eval.json"))
    print(json.dumps(j["eval"], indent=1)[:6000])
except Exception as e:
    print("eval ERR", e)
PY
