"""Stall-reason totals (and top stalled source lines) per kernel from
`ncu -i rep --page source --csv --print-source cuda,sass`; a kernel's rows are spread over
one block per inlined source file.  usage: ncu_stalls.py file.csv <kernel substring> [top]"""
import csv, sys, collections
path, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
csv.field_size_limit(1 << 30)
fn = None; fpath = None; hdr = None
agg = collections.Counter(); lines = collections.Counter(); lsrc = {}; lstall = collections.defaultdict(collections.Counter)
def ival(x):
    try: return int(x)
    except ValueError: return 0
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': fn = r[1]; continue
    if r[0] == 'Line No': hdr = r; names = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]; idx = {n: hdr.index(n) for n in names}; si = hdr.index('# Samples'); continue
    if r[0].isdigit() and hdr and len(r) == len(hdr) and fn and key in fn:
        k = (fpath, int(r[0])); lines[k] += ival(r[si]); lsrc[k] = r[1]
        for n in names:
            v = ival(r[idx[n]]); agg[n] += v; lstall[k][n[6:]] += v
tot = sum(lines.values()) or 1
print('total samples', tot)
for n, v in agg.most_common(12): print(f'  {n:26s} {v:8d} {100*v/tot:5.1f}%')
print('top lines by samples:')
for k, v in lines.most_common(top):
    print(f'  {100*v/tot:5.1f}% {k[0]}:{k[1]} {lsrc[k].strip()[:70]:70s} {lstall[k].most_common(2)}')
