#!/bin/bash
# round-2 GPU session 15: pair-kernel block size (register residency) at d=64 / d=128; parity of the capped general kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_pytest.log
tail -4 gpurun_out/r2n_pytest.log
for d in 64 128; do
  for pt in 256 128; do
    APR_PAIR_THREADS=$pt timeout 300 python bench.py --gpus 1 --steps 1024 --warmup 64 --dim $d --no-eval --no-variants --no-cpu > gpurun_out/r2n_bench_d${d}_pt$pt.json 2> gpurun_out/r2n_bench_d${d}_pt$pt.err
    python - $d $pt <<'PY'
import json, sys
d, pt = sys.argv[1:3]
try:
    j = json.load(open("gpurun_out/r2n_bench_d%s_pt%s.json" % (d, pt))); r = j["roofline"]
    print("PAIR d=%s pair_threads=%s value %.0fM ms %.4f kern %.4f frac %.3f whole %.3f" % (d, pt, j["value"]/1e6, j["ms_per_step"], r["ms_per_step_kernel"], r["frac"], r["whole_step_frac"]))
except Exception as e:
    print("PAIR", d, pt, "ERR", e)
PY
  done
done
