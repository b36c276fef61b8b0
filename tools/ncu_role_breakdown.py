"""Bucket the SASS-level sampling of an `ncu --set full --import-source on` report by code region (= warp role of the tc32 kernels):
  ncu -i REPORT.ncu-rep --page source --csv --print-source sass > src.csv;  python tools/ncu_role_breakdown.py src.csv [0x2000]
prints per region: samples, executed warp instructions, top stall reasons, instruction mix (profiles/r2_ncu_role_breakdown.txt)."""
import csv, sys
from collections import defaultdict
fn = sys.argv[1]; gran = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0x2000
rows = list(csv.reader(open(fn)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def f(r, c):
    try: return float(r[ix[c]])
    except: return 0.0
base = int(data[0][ix['Address']], 16)
b = defaultdict(lambda: defaultdict(float)); ex = defaultdict(float); ops = defaultdict(lambda: defaultdict(float))
for r in data:
    a = (int(r[ix['Address']], 16) - base) // gran
    for c in stall_cols: b[a][c] += f(r, c)
    b[a]['n'] += f(r, '# Samples'); ex[a] += f(r, 'Instructions Executed')
    src = r[ix['Source']].split()
    op = src[1] if src and src[0].startswith('@') and len(src) > 1 else (src[0] if src else '')
    for key in ('LDTM', 'STG', 'LDG', 'UTCHMMA', 'SHFL', 'UTMALDG', 'STS', 'LDS', 'SYNCS', 'F2FP', 'FADD', 'LDL', 'STL'):
        if op.startswith(key): ops[a][key] += f(r, 'Instructions Executed')
tot = sum(v['n'] for v in b.values())
for a in sorted(b):
    if b[a]['n'] < tot * 0.004: continue
    st = sorted(((b[a][c], c[6:]) for c in stall_cols), reverse=True)[:4]
    print("+%06x smp %6.0f (%4.1f%%) exec %9.0f | %s | %s" % (a * gran, b[a]['n'], 100 * b[a]['n'] / tot, ex[a], ", ".join("%s %.0f" % (c, v) for v, c in st if v > 0), " ".join("%s:%.0f" % kv for kv in ops[a].items())))
