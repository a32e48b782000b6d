#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_recommender.py tests/test_gpu_eval_tc.py -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_pytest.log
tail -25 gpurun_out/r2l_pytest.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-eval --no-variants --no-cpu > gpurun_out/r2l_bench_20.json 2> gpurun_out/r2l_bench_20.err
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2l_bench_20.json")); r = j["roofline"]
print("BENCH value %.0fM ms %.4f kern %.4f frac %.3f whole %.3f e2e %.0fM" % (j["value"]/1e6, j["ms_per_step"], r["ms_per_step_kernel"], r["frac"], r["whole_step_frac"], j["e2e"]["value"]/1e6))
PY
