#!/bin/bash
# round-2 ncu session (all ncu runs of one gpurun call): launch lists + full captures.  The .ncu-rep files are turned
# into CSV pages on the box and deleted there (gpurun_out/ is capped at 64 MiB); summaries go to profiles/.
set -x
mkdir -p gpurun_out
python tools/profile_step.py 128 20 > gpurun_out/r2p_plain_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2p_launches_step_B65536_d128.csv python tools/profile_step.py 128 20 > gpurun_out/r2p_ncu_step.log 2>&1
echo rc=$?; tail -3 gpurun_out/r2p_plain_step.log
full() {  # name, kernel regex, skip, count, command...
  name=$1; regex=$2; skip=$3; count=$4; shift 4
  "$@" > gpurun_out/${name}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $count -o /tmp/$name "$@" > gpurun_out/${name}_ncu.log 2>&1
  echo "rc=$?"
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2> /dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv 2> /dev/null | gzip -9 > gpurun_out/${name}_source.csv.gz
  ls -la /tmp/$name.ncu-rep gpurun_out/${name}_raw.csv gpurun_out/${name}_source.csv.gz
  rm -f /tmp/$name.ncu-rep
}
full r2p_step_full_d128 "fast_kernel|pair_kernel|general_stage_kernel" 20 10 python tools/profile_step.py 128 4
full r2p_step_full_d64 "fast_kernel|pair_kernel|general_stage_kernel" 20 5 python tools/profile_step.py 64 4
full r2p_prep_full_d128 "prep_" 0 12 python tools/profile_step.py 128 4
python tools/tcprof_topk.py 128 > gpurun_out/r2p_plain_tc.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2p_launches_tc_topk_16384x524288_d128.csv python tools/tcprof_topk.py 128 > gpurun_out/r2p_ncu_tc.log 2>&1
echo rc=$?; tail -2 gpurun_out/r2p_plain_tc.log
full r2p_tc_topk_full "tc_count_kernel|tc_topk|topk_" 6 6 python tools/tcprof_topk.py 128
du -sh gpurun_out
