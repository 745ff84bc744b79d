#!/usr/bin/env python
"""Top stall sites of an ncu report's SASS source page: tools/ncu_hot.py rep.ncu-rep [n] [stall_column]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
col = sys.argv[3] if len(sys.argv) > 3 else "# Samples"
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(float(r[ix["# Samples"]] or 0) for r in data)
def f(r, k):
    try: return float(r[ix[k]] or 0)
    except Exception: return 0.0
order = sorted(range(len(data)), key=lambda i: -f(data[i], col))[:n]
print(f"total samples {tot:.0f}; columns: idx samples% long_sb short_sb wait math mio instr_exec | sass")
for i in sorted(order):
    r = data[i]
    print(f"{i:6d} {100*f(r,'# Samples')/tot:5.2f}% lsb={f(r,'stall_long_sb'):5.0f} ssb={f(r,'stall_short_sb'):5.0f} wait={f(r,'stall_wait'):5.0f} math={f(r,'stall_math'):4.0f} exec={f(r,'Instructions Executed'):9.0f} | {r[ix['Source']].strip()[:90]}")
