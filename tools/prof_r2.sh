#!/bin/bash
# one gpurun call: plain runs first, then one ncu capture per kernel of interest; every report is summarised on the
# box (tools/ncu_summary.py) into gpurun_out/prof_r2/*.txt and deleted (the reports exceed what gpurun brings back)
set -x
mkdir -p gpurun_out/prof_r2
OUT=gpurun_out/prof_r2
N="ncu --set full --clock-control none"
cap() {  # cap <name> <kernel regex> <skip> <target...>
  name=$1; rx=$2; skip=$3; shift 3
  $N -k regex:$rx -s $skip -c 1 -o /tmp/$name python tools/prof_r2.py "$@" > /dev/null 2>&1
  python tools/ncu_summary.py kernel /tmp/$name.ncu-rep > $OUT/$name.txt 2>&1
  rm -f /tmp/$name.ncu-rep
}
python tools/prof_r2.py closest > $OUT/plain_closest.log 2>&1 && {
  cap hamming_r2_ncu_full k_hamming 1 closest
  cap select_row_r2_ncu_full k_select_row 1 closest
  cap consensus_r2_ncu_full k_consensus 1 closest
}
python tools/prof_r2.py ckpt > $OUT/plain_ckpt.log 2>&1 && {
  cap forward_ck_r2_ncu_full k_forward_ck 2 ckpt
  cap backward_ck_r2_ncu_full k_backward_ck 1 ckpt
}
python tools/prof_r2.py column > $OUT/plain_column.log 2>&1 && {
  cap column_r2_ncu_full k_column 2 column
  cap commit_all_r2_ncu_full k_commit_all 0 column
}
python tools/prof_r2.py upload > $OUT/plain_upload.log 2>&1 && {
  cap gather_r2_ncu_full k_gather 9 upload
}
python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --no-mcmc > $OUT/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file /tmp/r2_launches.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 0 --no-mcmc > /dev/null 2>&1
python tools/ncu_summary.py launches /tmp/r2_launches.csv > $OUT/launches_r2.txt 2>&1
ls -la $OUT
