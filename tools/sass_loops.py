#!/usr/bin/env python
"""Instruction mix of the loops of one kernel: tools/sass_loops.py LIB.so 'k_sweepILb0' [minlen] [lds_count]
(cuobjdump -sass; backward branches delimit loops).  Used to count issue slots / FP64 instructions per pair."""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 60
want_lds = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else None
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for l in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    a2i = {a: i for i, (a, _) in enumerate(ins)}
    print("==", name, len(ins), "instructions")
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in a2i:
            continue
        body = ins[a2i[tgt]:i + 1]
        if len(body) < minlen:
            continue
        c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0] for _, s in body)
        fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
        if want_lds is not None and c["LDS"] != want_lds:
            continue
        print(f"loop {tgt:#x}-{a:#x} n={len(body)} fp64={fp64} lds={c['LDS']} ldg={c['LDG']} :: " + " ".join(f"{k}:{v}" for k, v in c.most_common(16)))
        if "--dump" in sys.argv:
            for aa, s in body:
                print(f"    {aa:#07x}  {s}")
