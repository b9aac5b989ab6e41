"""Per-opcode stall-reason breakdown from an `ncu --page source --csv --print-source sass` dump."""
import csv
import re
import sys
from collections import Counter, defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc = hdr.index("Source")
reasons = ["stall_wait", "stall_mio", "stall_long_sb", "stall_short_sb", "stall_not_selected", "stall_dispatch",
           "stall_math", "stall_barrier", "stall_branch_resolving", "stall_no_inst", "stall_selected"]
idx = {r: hdr.index(r) for r in reasons}
by_op = defaultdict(Counter)
tot = Counter()
for r in rows[2:]:
    if len(r) <= max(idx.values()):
        continue
    toks = r[isrc].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = re.sub(r"\..*", "", op)
    for name, i in idx.items():
        v = int(r[i] or 0)
        by_op[op][name] += v
        tot[name] += v
alls = sum(tot.values())
print("totals:", {k: f"{100 * v / alls:.1f}%" for k, v in tot.most_common()})
print(f"{'op':10s} " + " ".join(f"{r[6:12]:>7s}" for r in reasons))
for op, c in sorted(by_op.items(), key=lambda kv: -sum(kv[1].values()))[:22]:
    print(f"{op:10s} " + " ".join(f"{100 * c[r] / alls:7.2f}" for r in reasons))
