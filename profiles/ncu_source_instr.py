"""Per-source-line executed warp-instruction counts from `ncu --page source --csv --print-source cuda,sass`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
fname = None; hdr = None; agg = collections.Counter(); src = {}
for r in rows:
    if not r:
        continue
    if r[0] in ('File Path', 'File Name'):
        fname = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = r; ci = hdr.index('Instructions Executed'); continue
    if hdr is None or len(r) <= ci or r[2] != '-':
        continue
    try:
        n = int(r[ci])
    except ValueError:
        continue
    agg[(fname, int(r[0]))] += n; src[(fname, int(r[0]))] = r[1].strip()[:90]
tot = sum(agg.values()); print('total warp-instr', tot)
for k, n in agg.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{100*n/tot:5.1f}% {k[0]}:{k[1]:<4d} {src[k]}")
