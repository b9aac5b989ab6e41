"""Summarise an `ncu --page source --csv --print-source sass` dump: executed warp-instructions per
opcode.  usage: python tools/sass_hist.py file.csv [divisor]"""
import csv
import re
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, total, samples = Counter(), 0, Counter()
for r in rows[2:]:
    if len(r) <= iex or not r[iex]:
        continue
    n = int(r[iex])
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = re.sub(r"\..*", "", op)
    ops[op] += n
    samples[op] += int(r[ismp] or 0)
    total += n
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print(f"total executed warp-instructions {total}  (/{div:g} = {total / div:.2f})")
for op, n in ops.most_common(32):
    print(f"{op:12s} {n:12d} {n / div:8.2f} {100 * n / total:6.2f}%  samples {samples[op]}")
