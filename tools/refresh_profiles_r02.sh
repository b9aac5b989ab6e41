#!/bin/bash
# GPU box, round 2: plain bench, the ncu launch list of the same command, then one `ncu --set full` capture per kernel
# (each only after its own command exited 0 without ncu).  Everything lands in gpurun_out/r02p_*; the summaries that are
# committed under profiles/ are made from these by tools/summarise_profiles_r02.py on the build machine.
set -x
O=gpurun_out
python bench.py --steps 5 --warmup 3 > $O/r02p_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02p_bench_launches.csv \
    python bench.py --steps 5 --warmup 3 > $O/r02p_ncu_launches.log 2>&1
cap() {  # cap <tag> <kernel regex> <skip> <script...>
  tag=$1; rx=$2; skip=$3; shift 3
  "$@" > $O/r02p_${tag}_plain.log 2>&1 || { echo "plain run of $tag failed"; return 1; }
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/r02p_${tag} "$@" > $O/r02p_${tag}_ncu.log 2>&1
  ncu -i /tmp/r02p_${tag}.ncu-rep --page raw --csv > $O/r02p_${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/r02p_${tag}.ncu-rep --page source --csv --print-source sass > $O/r02p_${tag}_sass.csv 2>/dev/null
}
cap stream_k4096_on step_stream 1 python tools/prof_step.py --launches 3
cap stream_k4096_off step_stream 1 python tools/prof_step.py --launches 3 --no-guidance --videos 32
cap stream_k2048_on step_stream 1 python tools/prof_step.py --launches 3 --codes 2048 --videos 32
cap stream_k1024_on step_stream 1 python tools/prof_step.py --launches 3 --codes 1024 --videos 64
cap head head_step_kernel 1 python tools/prof_head.py --launches 3
cap train train_stream_kernel 50 python tools/train_bench.py
ls -la $O/r02p_*
