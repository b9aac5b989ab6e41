"""Executed warp-instructions per SOURCE LINE: joins an `ncu --page source --csv --print-source sass` dump (per-SASS-
instruction executed counts) with the line table of the same kernel from `nvdisasm -g -c` of the shipped cubin.

    cuobjdump -xelf all libd3pm_b200.so && nvdisasm -g -c d3pm_api.sm_100a.cubin > all.sass
    python tools/sass_lines.py dump_sass.csv all.sass 'step_stream_kernelILi4ELb1ELb0' [divisor] [min_share]
"""
import csv
import re
import sys
from collections import Counter

dump, sass, pat = sys.argv[1], sys.argv[2], sys.argv[3]
div = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
min_share = float(sys.argv[5]) if len(sys.argv) > 5 else 0.004

lines, cur, on = [], None, False
for l in open(sass):
    if l.startswith(".text."):
        on = pat in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2).strip()))

rows = list(csv.reader(open(dump)))
hdr = rows[1]
iaddr, iex, ismp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
body = [r for r in rows[2:] if len(r) > iex and r[iex]]
base = int(body[0][iaddr], 16)
by_off = {off: (src, txt) for off, src, txt in lines}
per_line, per_line_s, total, miss = Counter(), Counter(), 0, 0
for r in body:
    off = int(r[iaddr], 16) - base
    n = int(r[iex])
    total += n
    if off not in by_off:
        miss += n
        continue
    per_line[by_off[off][0]] += n
    per_line_s[by_off[off][0]] += int(r[ismp] or 0)
print(f"{len(body)} SASS instructions in the dump, {len(lines)} in the cubin; executed {total} (/{div:g} = {total / div:.2f}); unmatched {miss}")
for (f, ln), n in sorted(per_line.items(), key=lambda kv: -kv[1]):
    if n / total < min_share:
        break
    print(f"{f}:{ln:<5d} {n:12d} {n / div:8.3f} {100 * n / total:6.2f}%  samples {per_line_s[(f, ln)]}")
