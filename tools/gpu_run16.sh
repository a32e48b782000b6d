#!/bin/bash
set -x
mkdir -p gpurun_out
for d in 64 128; do
  for pt in 64 128; do
    APR_PAIR_THREADS=$pt timeout 300 python bench.py --gpus 1 --steps 1024 --warmup 64 --dim $d --no-eval --no-variants --no-cpu > gpurun_out/r2o_bench_d${d}_pt$pt.json 2> gpurun_out/r2o_bench_d${d}_pt$pt.err
    python - $d $pt <<'PY'
import json, sys
d, pt = sys.argv[1:3]
try:
    j = json.load(open("gpurun_out/r2o_bench_d%s_pt%s.json" % (d, pt))); r = j["roofline"]
    print("PAIR d=%s pair_threads=%s value %.0fM ms %.4f kern %.4f frac %.3f whole %.3f" % (d, pt, j["value"]/1e6, j["ms_per_step"], r["ms_per_step_kernel"], r["frac"], r["whole_step_frac"]))
except Exception as e:
    print("PAIR", d, pt, "ERR", e)
PY
  done
done
APR_PAIR_THREADS=64 APR_PREP_BLOCKS=16 timeout 300 python bench.py --gpus 1 --steps 1024 --warmup 64 --dim 64 --no-eval --no-variants --no-cpu > gpurun_out/r2o_bench_d64_pt64_prep16.json 2>/dev/null
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2o_bench_d64_pt64_prep16.json")); r = j["roofline"]
print("PAIR d=64 pt=64 prep16 value %.0fM ms %.4f kern %.4f frac %.3f whole %.3f" % (j["value"]/1e6, j["ms_per_step"], r["ms_per_step_kernel"], r["frac"], r["whole_step_frac"]))
PY
