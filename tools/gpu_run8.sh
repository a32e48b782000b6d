#!/bin/bash
# round-2 GPU session 8: canary + loader tests, d=64 stream-priority experiment, the driver's bench command in full
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_canaries.py tests/test_gpu_loader.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_pytest.log
tail -25 gpurun_out/r2g_pytest.log
for d in 64 128; do
  for gp in 0 1; do
    APR_GEN_PRIO=$gp timeout 300 python bench.py --gpus 1 --steps 1024 --warmup 64 --dim $d --no-eval --no-variants --no-cpu > gpurun_out/r2g_bench_d${d}_genprio$gp.json 2> gpurun_out/r2g_bench_d${d}_genprio$gp.err
  done
done
python - <<'PY'
import json
for d in (64, 128):
    for gp in (0, 1):
        try:
            j = json.load(open("gpurun_out/r2g_bench_d%d_genprio%d.json" % (d, gp))); r = j["roofline"]
            print("PRIO d=%d gen_prio=%d value %.0fM ms %.4f kern %.4f frac %.3f whole %.3f" % (d, gp, j["value"]/1e6, j["ms_per_step"], r["ms_per_step_kernel"], r["frac"], r["whole_step_frac"]))
        except Exception as e:
            print("PRIO", d, gp, "ERR", e)
PY
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_full.json 2> gpurun_out/r2g_bench_full.err ) 2> gpurun_out/r2g_bench_full.time; echo rc=$?
cat gpurun_out/r2g_bench_full.time; tail -c 600 gpurun_out/r2g_bench_full.err
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_ref.json 2> gpurun_out/r2g_ref.err ) 2> gpurun_out/r2g_ref.time; cat gpurun_out/r2g_ref.time
python - <<'PY'
import json
j = json.load(open("gpurun_out/r2g_bench_full.json"))
print({k: (v if not isinstance(v, (dict, list)) else "...") for k, v in j.items()})
print("roofline", j["roofline"]); print("e2e", j["e2e"]); print("cpu", j.get("cpu_baseline")); print("graph", j.get("graph_cache"))
v = j.get("variants", {})
for k in ("dim_64", "dim_256"):
    print(k, v.get(k))
print("sweep", [(s["batch"], round(s["triples_per_s"]/1e6)) for s in v.get("batch_sweep_uniform", [])])
e = j.get("eval", {})
print({k: ((v2.get("users_per_s"), v2.get("ms")) if isinstance(v2, dict) else v2) for k, v2 in e.items()})
PY
