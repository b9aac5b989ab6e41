OLD=$PWD/tools/probes/libstep_old.so; NEW=$PWD/gif-synthesis-with-discrete-diffusion_b200/csrc/libd3pm_b200.so
stat() { python -c "
import sys,re
s=sys.stdin.read(); v=[float(x) for x in re.findall(r\"'([0-9.]+)'\", s)]; v=v[5:]; print('$1 mean %.4f min %.4f max %.4f'%(sum(v)/len(v), min(v), max(v)))"; }
for i in 1 2; do
echo "== old";     D3PM_B200_LIB=$OLD python tools/prof_step.py --launches 40 | stat on; D3PM_B200_LIB=$OLD python tools/prof_step.py --launches 40 --no-guidance --videos 32 | stat off32
echo "== new"; D3PM_B200_LIB=$NEW python tools/prof_step.py --launches 40 | stat on; D3PM_B200_LIB=$NEW python tools/prof_step.py --launches 40 --no-guidance --videos 32 | stat off32
done
