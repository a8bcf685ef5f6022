"""Attribute SASS instruction counts of one csm_stream_kernel instantiation to source regions.
usage: nvdisasm -g -c csm_stream.sm_100a.cubin > dis.txt; python tools/code_size.py dis.txt ILi1ELi4E"""
import bisect
import re
import sys
from collections import Counter

src = open('csm_hf_b200/csrc/csm_stream.inl').read().split('\n')
marks = [(1, 'top')]
for i, l in enumerate(src, 1):
    m = re.match(r'(?:__device__ __forceinline__|__global__).*?(\w+)\(', l)
    if m:
        marks.append((i, m.group(1)))
cnt, cur, fn = Counter(), None, None
for line in open(sys.argv[1]):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\.text\.(\S+):', line)
    if m:
        fn = m.group(1)
        continue
    if fn and sys.argv[2] in fn and re.match(r'\s+/\*[0-9a-f]+\*/\s+\S', line):
        cnt[cur] += 1
tot = sum(cnt.values())
print('total', tot, 'instr', tot * 16 // 1024, 'KB')
b = Counter()
for (f, l), n in cnt.items():
    if f not in ('csm_stream.cu', 'csm_stream.inl'):
        b[f] += n
        continue
    i = bisect.bisect_right([m[0] for m in marks], l) - 1
    b[marks[max(i, 0)][1]] += n
for k, v in b.most_common():
    print(f"{k:24s}{v:7d} {v * 16 / 1024:7.1f} KB")
