#!/bin/bash
# A/B of two library builds on ONE box at config 2 (burst: 40 launches; sustained: 1200 back-to-back launches)
stat() { python -c "
import sys,re
s=sys.stdin.read(); v=[float(x) for x in re.findall(r\"'([0-9.]+)'\", s)]; v=v[5:]; print('$1 mean %.4f min %.4f max %.4f n %d'%(sum(v)/len(v), min(v), max(v), len(v)))"; }
NEW=$PWD/gif-synthesis-with-discrete-diffusion_b200/csrc/libd3pm_b200.so
BASE=${BASE:-$PWD/tools/probes/libbase.so}
for i in 1 2 3; do
for lib in $BASE $NEW; do
echo "== $(basename $lib)"
D3PM_B200_LIB=$lib python tools/prof_step.py --launches 40 | stat k4096_on_burst; sleep 2
D3PM_B200_LIB=$lib python tools/prof_step.py --launches 1200 | stat k4096_on_sustained; sleep 3
done; done
