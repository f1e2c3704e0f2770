#!/usr/bin/env python3
"""Shared-memory wavefronts per LDS/STS flavour (and the instructions with excess wavefronts) from an ncu source-page export.
python tools/ncu_smem.py gpurun_out/full_X.source.csv.gz [...]"""
import csv, gzip, sys
from collections import defaultdict
def analyze(f):
    rows = list(csv.reader(gzip.open(f, 'rt')))
    hdr = next(i for i, r in enumerate(rows) if "Source" in r)
    I = {n: i for i, n in enumerate(rows[hdr])}
    agg = defaultdict(lambda: [0, 0, 0, 0])
    top = []
    for r in rows[hdr + 1:]:
        src = r[I['Source']]
        parts = src.split()
        if not parts: continue
        op = parts[1] if parts[0].startswith('@') and len(parts) > 1 else parts[0]
        if not (op.startswith('LDS') or op.startswith('STS')): continue
        f_ = lambda k: float(r[I[k]] or 0)
        a = agg[op]; a[0] += f_('Instructions Executed'); a[1] += f_('L1 Wavefronts Shared'); a[2] += f_('L1 Wavefronts Shared Ideal'); a[3] += f_('L1 Wavefronts Shared Excessive')
        if f_('L1 Wavefronts Shared Excessive') > 0:
            top.append((f_('L1 Wavefronts Shared Excessive'), r[I['Address']], src[:70], f_('Instructions Executed'), f_('L1 Wavefronts Shared')))
    print(f)
    for op, a in sorted(agg.items()):
        print(f"  {op:12s} instr {a[0]:12.0f} wavefronts {a[1]:12.0f} ideal {a[2]:12.0f} excessive {a[3]:12.0f}  wf/instr {a[1]/max(a[0],1):.2f}")
    top.sort(reverse=True)
    for t in top[:10]: print("    ", t)
for f in sys.argv[1:]: analyze(f)
