#!/bin/bash
# A/B on ONE box (boxes of the pool differ by several percent): every instantiation of the stream kernel with the
# baseline build (tools/probes/libbase.so) and the current one, alternating, 40 launches each (first 5 dropped).
stat() { python -c "
import sys,re
s=sys.stdin.read(); v=[float(x) for x in re.findall(r\"'([0-9.]+)'\", s)]; v=v[5:]; print('$1 mean %.4f min %.4f max %.4f'%(sum(v)/len(v), min(v), max(v)))"; }
NEW=$PWD/gif-synthesis-with-discrete-diffusion_b200/csrc/libd3pm_b200.so
BASE=${BASE:-$PWD/tools/probes/libbase.so}
for i in 1 2; do
for lib in $BASE $NEW; do
echo "== $(basename $lib)"
D3PM_B200_LIB=$lib python tools/prof_step.py --launches 40 | stat k4096_on; sleep 1
D3PM_B200_LIB=$lib python tools/prof_step.py --launches 40 --no-guidance --videos 32 | stat k4096_off_32v; sleep 1
D3PM_B200_LIB=$lib python tools/prof_step.py --launches 40 --codes 2048 --videos 32 | stat k2048_on_32v; sleep 1
D3PM_B200_LIB=$lib python tools/prof_step.py --launches 40 --codes 2048 --no-guidance --videos 64 | stat k2048_off_64v; sleep 1
D3PM_B200_LIB=$lib python tools/prof_step.py --launches 40 --codes 1024 --videos 64 | stat k1024_on_64v; sleep 1
done; done
