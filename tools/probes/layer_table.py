import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'fp32', round(d['native_fp32_ms'],3), 'tf32', round(d['native_tf32_ms'],3))
for prec in ('fp32', 'tf32'):
  if 'layers_' + prec not in d: continue
  print(' ', prec)
  for l in d['layers_' + prec][:4]+d['layers_' + prec][-3:]:
    print(f"  M={l['M']:7d} Cin={l['Cin']:4d} taps={l['taps']:2d} cls={l['classes']} N={l['Nout']:5d} nt={l['n_tile']} {l['ms']:.3f} ms {l['TFLOPs']:.0f} TF/s")
