#!/bin/bash
# multi-GPU bench line: tools/gpu_mg.sh N TAG
set -u
N=$1; TAG=$2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err
python - <<PY
import json
for l in open('gpurun_out/bench_${TAG}_${N}gpu.json'):
    try: d=json.loads(l)
    except Exception: continue
    print('n_gpus', d['n_gpus'], 'value', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
    s=d.get('strong_scaling'); print('strong', s and {k: s[k] for k in ('walkers_total','walkers_per_gpu','value','ms_per_step','shape')})
    w=d.get('widom'); print('widom', w and (w.get('insertions_per_s'), w.get('mu_ex_kcal_mol'), w.get('mu_ex_sigma'), w.get('insertions_total')))
    print('mixture', d.get('mixture') and d['mixture'].get('moves_per_s'))
PY
tail -3 gpurun_out/bench_${TAG}_${N}gpu.err
