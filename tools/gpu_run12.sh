#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2k_pytest_all.log 2>&1 ) 2> gpurun_out/r2k_pytest_all.time; echo "pytest rc=$?"
tail -6 gpurun_out/r2k_pytest_all.log; cat gpurun_out/r2k_pytest_all.time
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2k_smoke.log
for k in 1 2 3; do
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-eval --no-variants --no-cpu > gpurun_out/r2k_bench_20_$k.json 2> gpurun_out/r2k_bench_20_$k.err
done
timeout 300 python bench.py --gpus 1 --steps 2048 --warmup 128 --no-eval --no-variants --no-cpu > gpurun_out/r2k_bench_2048.json 2> gpurun_out/r2k_bench_2048.err
python - <<'PY'
import json
for f in ("r2k_bench_20_1", "r2k_bench_20_2", "r2k_bench_20_3", "r2k_bench_2048"):
    try:
        j = json.load(open("gpurun_out/%s.json" % f)); r = j["roofline"]
        print("BENCH %s value %.0fM ms %.4f kern %.4f frac %.3f whole %.3f e2e %.0fM graph %s" % (f, j["value"]/1e6, j["ms_per_step"], r["ms_per_step_kernel"], r["frac"], r["whole_step_frac"], j["e2e"]["value"]/1e6, j["graph_cache"]))
    except Exception as e:
        print("BENCH", f, "ERR", e)
PY
