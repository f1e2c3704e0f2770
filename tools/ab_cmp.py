#!/usr/bin/env python3
"""Compare the A/B sweeps written by tools/ab.sh (min of two runs each; prints rows that differ by > 1.5 %)."""
import json, sys
def load(f):
    d = {}
    for l in open(f):
        l = l.strip()
        if l.startswith("{"):
            r = json.loads(l); d[(r["kind"], r["n"], r["variant"])] = r["ms"]
    return d
A1, A2, B1, B2 = [load(f"gpurun_out/ab_{x}.jsonl") for x in ("A1", "A2", "B1", "B2")]
thr = float(sys.argv[1]) if len(sys.argv) > 1 else 0.015
for k in A1:
    if k in B1 and k in A2 and k in B2:
        a, b = min(A1[k], A2[k]), min(B1[k], B2[k])
        flag = "<<< A better" if a < b * (1 - thr) else (">>> B better" if b < a * (1 - thr) else "")
        if flag:
            print(k, A1[k], A2[k], "|", B1[k], B2[k], round(b / a, 3), flag)
