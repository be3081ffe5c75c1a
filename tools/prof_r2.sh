#!/bin/bash
# one gpurun call: plain runs first, then one ncu capture per kernel of interest (reports into gpurun_out/)
set -x
N="ncu --set full --clock-control none --import-source on"
python tools/prof_r2.py closest > gpurun_out/p_closest.log 2>&1 && {
  $N -k regex:k_hamming -s 1 -c 1 -o gpurun_out/r2_hamming python tools/prof_r2.py closest > /dev/null 2>&1
  $N -k regex:k_select_row -s 1 -c 1 -o gpurun_out/r2_select python tools/prof_r2.py closest > /dev/null 2>&1
  $N -k regex:k_consensus -s 1 -c 1 -o gpurun_out/r2_consensus python tools/prof_r2.py closest > /dev/null 2>&1
}
python tools/prof_r2.py ckpt > gpurun_out/p_ckpt.log 2>&1 && {
  $N -k regex:k_forward_ck -s 2 -c 1 -o gpurun_out/r2_forward_ck python tools/prof_r2.py ckpt > /dev/null 2>&1
  $N -k regex:k_backward_ck -s 1 -c 1 -o gpurun_out/r2_backward_ck python tools/prof_r2.py ckpt > /dev/null 2>&1
}
python tools/prof_r2.py column > gpurun_out/p_column.log 2>&1 && {
  $N -k regex:k_column -s 2 -c 1 -o gpurun_out/r2_column python tools/prof_r2.py column > /dev/null 2>&1
  $N -k regex:k_commit_all -c 1 -o gpurun_out/r2_commit python tools/prof_r2.py column > /dev/null 2>&1
}
python tools/prof_r2.py upload > gpurun_out/p_upload.log 2>&1 && {
  $N -k regex:k_gather -s 9 -c 1 -o gpurun_out/r2_gather python tools/prof_r2.py upload > /dev/null 2>&1
}
python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --no-mcmc > gpurun_out/p_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --no-mcmc > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv
