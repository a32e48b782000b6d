#!/bin/bash
# 1 GPU: final validation of the round -- whole GPU suite, smoke, the driver's bench command
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final_pytest.log
tail -4 gpurun_out/r2_final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/r2_final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench_20_5.json 2> gpurun_out/r2_final_bench_20_5.err; echo "bench rc=$?"
python - <<'PY'
import json
txt = open("gpurun_out/r2_final_bench_20_5.json").read()
j = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print("RES value %.1fM ms/step %.4f e2e %.1fM frac %.3f whole %.3f launches %s clocks %s" % (j["value"]/1e6, j["ms_per_step"], j["e2e"]["value"]/1e6, j["roofline"]["frac"], j["roofline"].get("whole_step_frac", -1), j["gpu_launches"], j["clocks"]))
print("RES cpu", j["cpu_baseline"])
print("RES eval", {k: (v.get("users_per_s") if isinstance(v, dict) else v) for k, v in j.get("eval", {}).items()})
print("RES variants", {k: (v.get("value"), v.get("roofline", {}).get("frac")) for k, v in j.get("variants", {}).items()})
PY
