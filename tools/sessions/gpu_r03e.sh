#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=r03e
OUT=gpurun_out/variants_$TAG.jsonl; : > $OUT
for ps in 13 9 5 1 15; do
  timeout 300 python bench.py --quick --steps 6 --warmup 3 --walkers 4736 --phase-sync $ps >> $OUT 2>> gpurun_out/variants_$TAG.err
done
python - <<'PY'
import json
for l in open('gpurun_out/variants_r03e.jsonl'):
    try:
        d = json.loads(l); print('sync %3s W %5s %8.3f M moves/s %7.2f ms C1 %.4f N %.1f' % (d.get('phase_sync'), d.get('walkers'), d['moves_per_s']/1e6, d['ms_per_step'], d.get('frac_c1', 0), d['loading'][1]))
    except Exception as e: print('bad', l[:80])
PY
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 1500 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
