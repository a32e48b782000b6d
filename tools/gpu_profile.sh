#!/bin/bash
# round-2 ncu session (all ncu runs of one gpurun call): launch lists + full captures; summaries go to profiles/
set -x
mkdir -p gpurun_out
python tools/profile_step.py 128 20 > gpurun_out/r2p_plain_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2p_launches_step_B65536_d128.csv python tools/profile_step.py 128 20 > gpurun_out/r2p_ncu_step.log 2>&1
echo rc=$?; tail -3 gpurun_out/r2p_plain_step.log
python tools/profile_step.py 128 4 > gpurun_out/r2p_plain_step4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fast_kernel|pair_kernel|general_stage_kernel" -s 20 -c 20 -o gpurun_out/r2p_step_full_d128 python tools/profile_step.py 128 4 > gpurun_out/r2p_ncu_step_full.log 2>&1
echo rc=$?
python tools/profile_step.py 64 4 > gpurun_out/r2p_plain_step4_d64.log 2>&1 &&
ncu --set full --clock-control none -k regex:"fast_kernel|pair_kernel|general_stage_kernel" -s 20 -c 20 -o gpurun_out/r2p_step_full_d64 python tools/profile_step.py 64 4 > gpurun_out/r2p_ncu_step_full_d64.log 2>&1
echo rc=$?
python tools/tcprof_topk.py 128 > gpurun_out/r2p_plain_tc.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2p_launches_tc_topk_16384x524288_d128.csv python tools/tcprof_topk.py 128 > gpurun_out/r2p_ncu_tc.log 2>&1
echo rc=$?; tail -2 gpurun_out/r2p_plain_tc.log
ncu --set full --clock-control none --import-source on -k regex:"tc_count_kernel|tc_topk|topk_" -s 6 -c 6 -o gpurun_out/r2p_tc_topk_full python tools/tcprof_topk.py 128 > gpurun_out/r2p_ncu_tc_full.log 2>&1
echo rc=$?
ls -la gpurun_out/*.ncu-rep
