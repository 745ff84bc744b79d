#!/usr/bin/env python
"""Warp-stall samples of an ncu report by source FUNCTION (or by line), using the source the report itself carries
(--import-source on), so the build that was profiled need not exist any more:
  ncu -i rep --page source --csv --print-source cuda      > cuda.csv
  ncu -i rep --page source --csv --print-source cuda,sass > cs.csv
  tools/ncu_attrib_src.py cuda.csv cs.csv [n_lines]   (n_lines given: top source lines instead of functions)"""
import csv,re,collections,bisect,sys
cuda_csv, cs_csv = sys.argv[1], sys.argv[2]
byline = len(sys.argv) > 3
full={}
for r in csv.reader(open(cuda_csv)):
    if len(r)==2 and r[0].isdigit(): full[int(r[0])]=r[1]
rows=list(csv.reader(open(cs_csv)))
hdr=None; lr=[]; fname=None
for r in rows:
    if len(r)==2 and r[0]=="File Path": fname=r[1]; continue
    if r and r[0]=="Line No":
        hdr=r; ix={}
        for i,h in enumerate(r): ix.setdefault(h,i)
        continue
    if hdr and len(r)==len(hdr) and r[0].isdigit() and r[2]=="-" and fname and fname.endswith("mgpu_kernels.cuh"): lr.append(r)
starts=[]
for ln in sorted(full):
    s=full[ln]
    if re.match(r'^(__device__|__global__|static __device__)', s):
        t=s if '(' in s else full.get(ln+1,'')
        m=re.search(r'(\w+)\s*\(', t)
        m2=re.search(r'\)\s*(\w+)\s*\(', s)
        name=(m2.group(1) if ('__launch_bounds__' in s and m2) else (m.group(1) if m else s[:30]))
        starts.append((ln,name))
    m3=re.match(r'^\s+__device__ __forceinline__ \w+ (load|fetch|block|run|run_guest|n_charged)\(', s)
    if m3: starts.append((ln,'HostPass::'+m3.group(1)))
starts.sort(); keys=[s[0] for s in starts]
def fn(ln):
    i=bisect.bisect_right(keys,ln)-1
    return starts[i][1] if i>=0 else '?'
def f(r,k):
    try: return float(r[ix[k]] or 0)
    except: return 0.0
ST=['stall_long_sb','stall_wait','stall_short_sb','stall_math','stall_no_inst','stall_barrier','stall_not_selected','stall_selected','stall_branch_resolving','stall_dispatch']
agg=collections.defaultdict(lambda:[0.0]*(len(ST)+2)); tot=0
for r in lr:
    ln=int(r[0]); k=(ln, full.get(ln,'').strip()[:90]) if byline else fn(ln); a=agg[k]; s=f(r,'# Samples'); a[0]+=s; a[1]+=f(r,'Instructions Executed'); tot+=s
    for i,kk in enumerate(ST): a[2+i]+=f(r,kk)
print('samples%  exec(M) | '+' '.join(k.replace('stall_','')[:8].rjust(8) for k in ST))
for k,a in sorted(agg.items(), key=lambda x:-x[1][0])[:int(sys.argv[3]) if byline else 30]:
    print(f"{100*a[0]/tot:5.1f}% {a[1]/1e6:8.1f} | "+' '.join(f"{100*v/max(a[0],1):8.0f}" for v in a[2:])+f" | {k}")
