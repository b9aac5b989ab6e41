stat() { python -c "
import sys,re
s=sys.stdin.read(); v=[float(x) for x in re.findall(r\"'([0-9.]+)'\", s)]; v=v[5:]; print('$1 mean %.4f min %.4f max %.4f'%(sum(v)/len(v), min(v), max(v)))"; }
for i in 1 2; do
for v in ${VARIANTS:-base t1}; do
lib=$PWD/tools/probes/lib$v.so
echo "== $v"; D3PM_B200_LIB=$lib python tools/prof_step.py --launches 40 | stat on; sleep 1
D3PM_B200_LIB=$lib python tools/prof_step.py --launches 40 --no-guidance --videos 32 | stat off32; sleep 1
D3PM_B200_LIB=$lib CHUNKS=10 python tools/probes/drift.py | awk '/chunk/{s+=$3; n++} END{printf " sustained mean %.4f\n", s/n}'
sleep 2
done; done
