#!/bin/bash
# round-2 GPU session 1: parity tests, then the bench at the driver's settings and at the long settings
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-eval --no-variants --no-cpu > gpurun_out/r2a_bench_20_5.json 2> gpurun_out/r2a_bench_20_5.err; echo rc=$?
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-eval --no-variants --no-cpu > gpurun_out/r2a_bench_20_5_b.json 2>> gpurun_out/r2a_bench_20_5.err; echo rc=$?
timeout 600 python bench.py --gpus 1 --steps 2048 --warmup 128 --no-eval --no-variants --no-cpu > gpurun_out/r2a_bench_2048.json 2> gpurun_out/r2a_bench_2048.err; echo rc=$?
APR_GRAPH=0 timeout 600 python bench.py --gpus 1 --steps 2048 --warmup 128 --no-eval --no-variants --no-cpu > gpurun_out/r2a_bench_2048_nograph.json 2> gpurun_out/r2a_bench_2048_nograph.err; echo rc=$?
timeout 600 python bench.py --gpus 1 --steps 2048 --warmup 128 --dim 64 --no-eval --no-variants --no-cpu > gpurun_out/r2a_bench_2048_d64.json 2> gpurun_out/r2a_bench_2048_d64.err; echo rc=$?
for b in 4096 16384; do
  timeout 300 python bench.py --gpus 1 --steps 2048 --warmup 128 --batch $b --no-eval --no-variants --no-cpu > gpurun_out/r2a_bench_B${b}.json 2> gpurun_out/r2a_bench_B${b}.err
  APR_GRAPH_MIN_BATCH=1024 timeout 300 python bench.py --gpus 1 --steps 2048 --warmup 128 --batch $b --no-eval --no-variants --no-cpu > gpurun_out/r2a_bench_B${b}_graph.json 2> gpurun_out/r2a_bench_B${b}_graph.err
done
head -c 600 gpurun_out/r2a_bench_20_5.json; echo; head -c 600 gpurun_out/r2a_bench_2048.json; echo
