#!/bin/bash
# GPU box: plain bench, then the ncu launch list of the same command, then one full capture of the stream kernel.
set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_plain_r1b.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv \
    python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_launches_r1b.log 2>&1
python tools/prof_step.py --launches 3 > gpurun_out/prof_plain_r1b.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:step_stream -s 1 -c 1 -f -o gpurun_out/prof_r1b_stream \
    python tools/prof_step.py --launches 3 > gpurun_out/ncu_r1b.log 2>&1
tail -2 gpurun_out/ncu_r1b.log
