#!/bin/bash
# round-2 GPU session 11: the whole -m gpu suite in one process (as the driver runs it), smoke, the driver's bench command
set -x
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2j_pytest_all.log 2>&1 ) 2> gpurun_out/r2j_pytest_all.time; echo "pytest rc=$?"
tail -6 gpurun_out/r2j_pytest_all.log; cat gpurun_out/r2j_pytest_all.time
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2j_smoke.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err ) 2> gpurun_out/r2j_bench.time; echo "bench rc=$?"; cat gpurun_out/r2j_bench.time
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2j_bench.json")); r = j["roofline"]
print("BENCH value %.0fM ms %.4f kern %.4f frac %.3f whole %.3f e2e %.0fM traffic %s graph %s" % (j["value"]/1e6, j["ms_per_step"], r["ms_per_step_kernel"], r["frac"], r["whole_step_frac"], j["e2e"]["value"]/1e6, r["traffic"], j["graph_cache"]))
PY
