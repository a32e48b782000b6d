#!/bin/bash
# usage: tools/gpu_sanitize.sh <memcheck|racecheck|synccheck|initcheck>   (ONE tool per gpurun call: B200_PROFILING.md)
tool=$1
set -x
mkdir -p gpurun_out
timeout 600 python tools/sanitize_case.py > gpurun_out/r2s_plain_$tool.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/r2s_plain_$tool.log; exit 1; }
tail -4 gpurun_out/r2s_plain_$tool.log
timeout 2400 compute-sanitizer --tool $tool --print-limit 50 --log-file gpurun_out/r2s_$tool.log python tools/sanitize_case.py > gpurun_out/r2s_run_$tool.log 2>&1
echo "sanitizer rc=$?"
tail -5 gpurun_out/r2s_run_$tool.log
tail -15 gpurun_out/r2s_$tool.log
