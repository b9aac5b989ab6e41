stat() { python -c "
import sys,re
s=sys.stdin.read(); v=[float(x) for x in re.findall(r\"'([0-9.]+)'\", s)]; v=v[5:]; print('$1 mean %.4f min %.4f max %.4f'%(sum(v)/len(v), min(v), max(v)))"; }
NEW=$PWD/gif-synthesis-with-discrete-diffusion_b200/csrc/libd3pm_b200.so
for i in 1 2; do
for lib in $PWD/tools/probes/libbase.so $NEW; do
echo "== $lib"; D3PM_B200_LIB=$lib python tools/prof_step.py --launches 40 | stat on; sleep 1
D3PM_B200_LIB=$lib python tools/prof_step.py --launches 40 --no-guidance --videos 32 | stat off32; sleep 1
done; done
python tools/probes/purity_time.py
