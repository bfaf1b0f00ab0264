"""Aggregate executed warp-instructions per SASS opcode from `ncu --page source --csv --print-source sass`."""
import csv, sys, collections
kern = None; hdr = None
agg = {}
for r in csv.reader(open(sys.argv[1])):
    if not r: continue
    if r[0] == 'Kernel Name':
        kern = r[1]; agg[kern] = collections.Counter(); continue
    if r[0] == 'Address':
        hdr = r; ie = hdr.index('Instructions Executed'); continue
    if kern is None or hdr is None or len(r) <= ie: continue
    toks = r[1].split()
    op = toks[0] if not toks[0].startswith('@') else toks[1]
    op = op.split('.')[0]
    try: agg[kern][op] += int(r[ie])
    except ValueError: pass
for k, c in agg.items():
    tot = sum(c.values())
    print('====', k[:70], 'total warp-instr', tot)
    for op, n in c.most_common(16):
        print(f'   {op:10s} {n:12d} {100*n/tot:5.1f}%')
