"""Executed warp-instructions per source line of one kernel from `ncu -i rep --page source --csv
--print-source cuda,sass`.  usage: ncu_lines.py file.csv <kernel substring> [top]"""
import csv, collections, sys
csv.field_size_limit(1 << 30)
path, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fn = None; fpath = None; hdr = None
agg = collections.Counter(); src = {}
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': fn = r[1]; continue
    if r[0] == 'Line No': hdr = r; ii = hdr.index('Instructions Executed'); continue
    if r[0].isdigit() and hdr and len(r) == len(hdr) and fn and key in fn:
        try: n = int(r[ii])
        except ValueError: continue
        agg[(fpath, int(r[0]))] += n; src[(fpath, int(r[0]))] = r[1].strip()
tot = sum(agg.values()) or 1
print('total warp-instructions', tot)
for k, v in agg.most_common(top): print(f'{100*v/tot:5.1f}% {k[0]}:{k[1]} {src[k][:100]}')
