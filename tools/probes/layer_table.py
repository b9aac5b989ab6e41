import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], 'fp32', round(d['native_fp32_ms'],3), 'tf32', round(d['native_tf32_ms'],3))
for l in d['layers_fp32'][:4]+d['layers_fp32'][-3:]:
    print(f"  M={l['M']:7d} Cin={l['Cin']:4d} taps={l['taps']:2d} cls={l['classes']} N={l['Nout']:5d} nt={l['n_tile']} {l['ms']:.3f} ms {l['TFLOPs']:.0f} TF/s")
