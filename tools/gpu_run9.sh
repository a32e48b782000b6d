#!/bin/bash
# round-2 GPU session 9: N4 test, stream-priority masks at d=64 / d=128
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_recommender.py -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_pytest.log
tail -15 gpurun_out/r2h_pytest.log
for d in 64 128; do
  for gp in 4 5 1; do
    APR_GEN_PRIO=$gp timeout 300 python bench.py --gpus 1 --steps 1024 --warmup 64 --dim $d --no-eval --no-variants --no-cpu > gpurun_out/r2h_bench_d${d}_mask$gp.json 2> gpurun_out/r2h_bench_d${d}_mask$gp.err
    python - $d $gp <<'PY'
import json, sys
d, gp = sys.argv[1:3]
try:
    j = json.load(open("gpurun_out/r2h_bench_d%s_mask%s.json" % (d, gp))); r = j["roofline"]
    print("PRIO d=%s mask=%s value %.0fM ms %.4f kern %.4f frac %.3f whole %.3f" % (d, gp, j["value"]/1e6, j["ms_per_step"], r["ms_per_step_kernel"], r["frac"], r["whole_step_frac"]))
except Exception as e:
    print("PRIO", d, gp, "ERR", e)
PY
  done
done
