#!/bin/bash
# GPU box: one `ncu --set full` capture per instantiation of step_stream_kernel, exported as CSV (raw page, SASS page with
# executed counts, CUDA-source page) under gpurun_out/<tag>_*; the .ncu-rep itself is not kept (size).
#   tools/capture_stream.sh <tag> [extra args of tools/prof_step.py ...]
tag=$1; shift
python tools/prof_step.py --launches 3 "$@" > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:step_stream -s 1 -c 1 -f -o /tmp/${tag} \
    python tools/prof_step.py --launches 3 "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i /tmp/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
ncu -i /tmp/${tag}.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_sass.csv 2>/dev/null
ncu -i /tmp/${tag}.ncu-rep --page source --csv --print-source cuda > gpurun_out/${tag}_cuda.csv 2>/dev/null
ls -la /tmp/${tag}.ncu-rep gpurun_out/${tag}_*
