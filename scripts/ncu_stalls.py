"""Stall-reason totals and top stalled source lines for one kernel from
`ncu --page source --csv --print-source cuda,sass`.  usage: ncu_stalls.py file.csv <kernel substring>"""
import csv, sys, collections
path, key = sys.argv[1], sys.argv[2]
fn = None; fpath = None; hdr = None; rows = []
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': fn = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0].isdigit() and hdr and len(r) == len(hdr) and fn and key in fn:
        rows.append((fpath, int(r[0]), r[1], r))
si = hdr.index('# Samples')
def ival(x):
    try: return int(x)
    except ValueError: return 0
tot = sum(ival(r[3][si]) for r in rows) or 1
names = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
idx = {n: hdr.index(n) for n in names}
agg = collections.Counter()
for f, l, s, r in rows:
    for n in names: agg[n] += ival(r[idx[n]])
print('total samples', tot)
for n, v in agg.most_common(10): print(f'  {n:26s} {v:8d} {100*v/tot:5.1f}%')
print('top lines by samples:')
for f, l, s, r in sorted(rows, key=lambda x: -ival(x[3][si]))[:22]:
    st = sorted(((ival(r[idx[n]]), n[6:]) for n in names), reverse=True)[:2]
    print(f'  {100*ival(r[si])/tot:5.1f}% {f}:{l} {s[:64]:64s} {st}')
