#!/usr/bin/env python
"""Attribute the warp-stall samples of an ncu report to source functions / lines.
usage: tools/ncu_attrib.py source_page.csv disasm_with_lineinfo.txt kernel_symbol_prefix source_file [lines]
  source_page.csv : ncu -i rep --page source --csv
  disasm          : nvdisasm --print-line-info cubin   (same build as the profiled .so)"""
import collections, csv, re, sys
csvf, disf, sym, srcf = sys.argv[1:5]
by_line = len(sys.argv) > 5
lines = open(disf).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + sym))
cur = None; seq = []
for l in lines[start + 1:]:
    if (l.startswith('.text.') or l.startswith('.section')) and seq: break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): seq.append(cur)
rows = list(csv.reader(open(csvf)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
assert len(data) == len(seq), (len(data), len(seq))
src = open(srcf).read().split('\n')
base = srcf.split('/')[-1]
def f(r, k):
    try: return float(r[ix[k]] or 0)
    except Exception: return 0.0
def func_of(line):
    for j in range(line - 1, 0, -1):
        s = src[j - 1]
        if re.match(r'^(template|__device__|__global__|static|__host__)', s) or re.match(r'^\s+__device__ __forceinline__ void (load|fetch|block|run|run_guest)', s):
            t = s if '(' in s else src[j]
            m = re.search(r'(\w+)\s*\(', t)
            return m.group(1) if m else s[:30]
    return '?'
STALLS = ['stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_math', 'stall_no_inst', 'stall_barrier', 'stall_not_selected', 'stall_branch_resolving', 'stall_dispatch', 'stall_lg', 'stall_mio']
agg = collections.defaultdict(lambda: [0.0] * (len(STALLS) + 2))
tot = 0.0
for r, c in zip(data, seq):
    if c and c[0] == base: key = (c[1], src[c[1] - 1].strip()[:70]) if by_line else func_of(c[1])
    else: key = c[0] if c else '?'
    a = agg[key]
    s = f(r, '# Samples'); a[0] += s; a[1] += f(r, 'Instructions Executed'); tot += s
    for i, k in enumerate(STALLS): a[2 + i] += f(r, k)
print('samples%  exec(M) | ' + ' '.join(k.replace('stall_', '')[:8] for k in STALLS))
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:int(sys.argv[5]) if by_line else 30]:
    print(f"{100 * a[0] / tot:5.1f}% {a[1] / 1e6:8.1f} | " + ' '.join(f"{100 * v / max(a[0], 1):8.0f}" for v in a[2:]) + f" | {k}")
