#!/bin/bash
# round-2 GPU session 10: resident-block sweeps of the single-GPU step (whole step, index preparation included)
set -x
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --gpus 1 --steps 1024 --warmup 64 --no-eval --no-variants --no-cpu > gpurun_out/r2i_$tag.json 2> gpurun_out/r2i_$tag.err
  python - $tag <<'PY'
import json, sys
try:
    j = json.load(open("gpurun_out/r2i_%s.json" % sys.argv[1])); r = j["roofline"]
    print("SWEEP %-14s value %.0fM ms %.4f kern %.4f frac %.3f whole %.3f e2e %.0fM" % (sys.argv[1], j["value"]/1e6, j["ms_per_step"], r["ms_per_step_kernel"], r["frac"], r["whole_step_frac"], j["e2e"]["value"]/1e6))
except Exception as e:
    print("SWEEP", sys.argv[1], "ERR", e)
PY
}
run base APR_X=0
run prep2 APR_PREP_BLOCKS=2
run prep4 APR_PREP_BLOCKS=4
run prep8 APR_PREP_BLOCKS=8
run prep32 APR_PREP_BLOCKS=32
run fast3 APR_FAST_BLOCKS=3
run fast4 APR_FAST_BLOCKS=4
run gen1 APR_GEN_BLOCKS=1
run gen1_prep4 APR_GEN_BLOCKS=1 APR_PREP_BLOCKS=4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2i_smoke.log
