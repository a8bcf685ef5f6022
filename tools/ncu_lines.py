"""Join an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g -c line info
and print the source lines with the most warp-stall samples.
usage: python tools/ncu_lines.py src.csv dis.txt KERNEL_SUBSTR [topN]"""
import csv
import re
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
isrc = hdr.index("Source")
sass = [(int(r[ia], 16), int(r[isamp] or 0), int(r[iex] or 0), r[isrc].strip()) for r in rows[2:] if len(r) > isamp and r[ia].startswith("0x")]
base = sass[0][0]
lines, cur, fn = {}, None, None
for line in open(sys.argv[2]):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\.text\.(\S+):', line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(\S.*)', line)
    if fn and sys.argv[3] in fn and m:
        lines[int(m.group(1), 16)] = cur
samp, exe = Counter(), Counter()
tot = 0
for a, s, e, txt in sass:
    k = lines.get(a - base)
    samp[k] += s
    exe[k] += e
    tot += s
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = open('csm_hf_b200/csrc/csm_stream.inl').read().split('\n')
bat = open('csm_hf_b200/csrc/csm_batch.inl').read().split('\n')
com = open('csm_hf_b200/csrc/csm_common.cuh').read().split('\n')
print("total samples", tot)
for k, s in samp.most_common(top):
    text = ""
    if k and k[0] in ('csm_stream.cu', 'csm_stream.inl'):
        text = src[k[1] - 1].strip()
    elif k and k[0] == 'csm_batch.inl':
        text = bat[k[1] - 1].strip()
    elif k and k[0] == 'csm_common.cuh':
        text = com[k[1] - 1].strip()
    print(f"{100 * s / tot:5.1f}% {s:7d} exec {exe[k]:10d}  {k}  {text[:110]}")
