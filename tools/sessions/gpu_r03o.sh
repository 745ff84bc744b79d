#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=r03o
OUT=gpurun_out/variants_$TAG.jsonl; : > $OUT
V=$PWD/maniac-mc.github.io_b200/variants
q() { lib=$1; shift; if [ "$lib" = default ]; then unset MANIAC_GPU_LIB; else export MANIAC_GPU_LIB=$V/libmaniac_gpu_$lib.so; fi; timeout 300 python bench.py --quick --steps 6 --warmup 3 "$@" >> $OUT 2>> gpurun_out/variants_$TAG.err; }
q default --walkers 4736
q pfg --walkers 4736
q default --walkers 4736 --loading 128
q pfg --walkers 4736 --loading 128
unset MANIAC_GPU_LIB
python - <<'PY'
import json
for l in open('gpurun_out/variants_r03o.jsonl'):
    try:
        d = json.loads(l); print('%-12s sync %3s W %5s %8.3f M moves/s %7.2f ms C1 %.4f N %.1f' % (d['lib'][-12:], d.get('phase_sync'), d.get('walkers'), d['moves_per_s']/1e6, d['ms_per_step'], d.get('frac_c1', 0), d['loading'][1]))
    except Exception as e: print('bad', l[:80])
PY
