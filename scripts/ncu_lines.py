"""Top source lines by executed warp-instructions from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys, collections
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
fn = None; fpath = None
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
src = {}
for r in csv.reader(open(sys.argv[1])):
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': fn = r[1][:60]; continue
    if r[0] == 'Line No': continue
    if r[0].isdigit() and len(r) > 8:
        try: n = int(r[7])
        except ValueError: continue
        agg[fn][(fpath, int(r[0]))] += n; tot[fn] += n; src[(fpath, int(r[0]))] = r[1]
for fn, c in agg.items():
    print('====', fn, 'total', tot[fn])
    for (f, l), n in c.most_common(top):
        print(f'  {100*n/tot[fn]:5.1f}%  {n:11d}  {f}:{l}  {src[(f,l)][:90]}')
