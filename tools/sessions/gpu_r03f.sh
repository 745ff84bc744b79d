#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=r03f
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --tb=short -x -k "sweep and not 10k" > gpurun_out/pytest_$TAG.log 2>&1
tail -3 gpurun_out/pytest_$TAG.log
OUT=gpurun_out/variants_$TAG.jsonl; : > $OUT
q() { timeout 300 python bench.py --quick --steps 6 --warmup 3 "$@" >> $OUT 2>> gpurun_out/variants_$TAG.err; }
q --walkers 4736
q --walkers 4736 --loading 16
q --walkers 4736 --loading 16 --phase-sync 13
q --walkers 4736 --loading 128
q --walkers 4736 --loading 128 --phase-sync 13
q --walkers 512
q --walkers 512 --phase-sync 5
q --walkers 1024
q --walkers 1024 --phase-sync 5
python - <<'PY'
import json
for l in open('gpurun_out/variants_r03f.jsonl'):
    try:
        d = json.loads(l); print('sync %3s W %5s %8.3f M moves/s %7.2f ms C1 %.4f N %.1f' % (d.get('phase_sync'), d.get('walkers'), d['moves_per_s']/1e6, d['ms_per_step'], d.get('frac_c1', 0), d['loading'][1]))
    except Exception as e: print('bad', l[:80])
PY
M=gpurun_out/mixture_$TAG.log; : > $M
for ps in 1 0 5 13 21; do timeout 200 python tools/mixture_probe.py 2368 32 $ps >> $M 2>&1; done
cat $M
