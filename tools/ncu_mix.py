"""Instruction mix / stall summary from `ncu --page source --csv` output."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
op = collections.Counter(); samp = collections.Counter(); tot = 0
for r in data:
    if len(r) < len(hdr): continue
    src = r[ix['Source']]; n = int(float(r[ix['Instructions Executed']] or 0)); s = int(float(r[ix['# Samples']] or 0))
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    o = m.group(2).split('.')[0] if m else '?'
    op[o] += n; samp[o] += s; tot += n
print('total warp-instructions', tot, 'SASS lines', len(data))
for o, n in op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f'{o:12s} {n:12d} {100*n/tot:5.1f}%  samples {samp[o]}')
