"""Rank CUDA source lines of an ncu --import-source report by warp-stall samples.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv; python tools/ncu_lines.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or r[2] != '-':
        continue
    try:
        s = int(r[hdr.index('# Samples')])
    except ValueError:
        continue
    out.append((s, cur, r))
tot = sum(s for s, _, _ in out)
print('total samples', tot)
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
ie = hdr.index('Instructions Executed')
out.sort(key=lambda x: -x[0])
for s, f, r in out[:top]:
    st = sorted([(int(r[i] or 0), hdr[i][6:]) for i in stall], reverse=True)[:3]
    print(f"{s:7d} {100 * s / tot:5.1f}% {f}:{r[0]:>4s} inst={int(r[ie]):>10d} {r[1].strip()[:84]:84s} {st}")
