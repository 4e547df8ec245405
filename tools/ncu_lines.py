"""Aggregate an ncu SASS-page CSV per source line using nvdisasm -g line info.
usage: ncu_lines.py <sass.csv> <nvdisasm -g output> <kernel substring> [top]"""
import collections
import csv
import re
import sys

sass_csv, disasm, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(disasm).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('//--------------------- .text.') and kern in l)
addr2line, cur = {}, None
for l in lines[start + 1:]:
    if l.startswith('//--------------------- .') and 'text' in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hi[0]]
col = {n: h.index(n) for n in ['Address', '# Samples', 'Instructions Executed', 'stall_long_sb', 'stall_wait',
                               'stall_short_sb', 'stall_branch_resolving']}
end = hi[1] - 1 if len(hi) > 1 else len(rows)
agg = collections.defaultdict(lambda: [0] * 6)
tot = totx = 0
base = None
for r in rows[hi[0] + 1:end]:
    if len(r) < len(h):
        continue
    a = int(r[col['Address']], 16)
    base = a if base is None else base
    ln = addr2line.get(a - base)
    v = agg[ln]
    vals = [int(r[col[n]] or 0) for n in ['# Samples', 'Instructions Executed', 'stall_long_sb', 'stall_wait',
                                           'stall_short_sb', 'stall_branch_resolving']]
    for i, x in enumerate(vals):
        v[i] += x
    tot += vals[0]
    totx += vals[1]
print("total samples", tot, "total warp-instructions", totx)
src = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = str(k)
    if k and (k[0].endswith('.cu') or k[0].endswith('.cuh')):
        try:
            src.setdefault(k[0], open('/root/repo/hnsw_slim_b200/csrc/' + k[0]).read().split('\n'))
            text = src[k[0]][k[1] - 1].strip()[:64]
        except OSError:
            pass
    print(f"{str(k):34s} samp {100*v[0]/tot:5.1f}% inst {100*v[1]/totx:5.1f}% long {v[2]:6d} wait {v[3]:6d} "
          f"short {v[4]:6d} br {v[5]:6d} | {text}")
