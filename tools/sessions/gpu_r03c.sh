#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=r03c
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --tb=short -x -k "automatic_sweep_shapes or team2" > gpurun_out/pytest_$TAG.log 2>&1
tail -5 gpurun_out/pytest_$TAG.log
OUT=gpurun_out/variants_$TAG.jsonl; : > $OUT
V=$PWD/maniac-mc.github.io_b200/variants
run() { # lib walkers extra...
  lib=$1; w=$2; shift; shift
  if [ "$lib" = default ]; then unset MANIAC_GPU_LIB; else export MANIAC_GPU_LIB=$V/libmaniac_gpu_$lib.so; fi
  timeout 300 python bench.py --quick --steps 6 --warmup 3 --walkers $w "$@" >> $OUT 2>> gpurun_out/variants_$TAG.err
}
for rep in 1 2; do
  run default 4736; run ag2 4736; run ag1 4736
done
for w in 512 1024 2048 4096; do run base $w; run default $w; done
run default 512 --phase-sync 0
run default 1024 --phase-sync 0
run ag2 2048
run ag2 4096
unset MANIAC_GPU_LIB
python - <<'PY'
import json
for l in open('gpurun_out/variants_r03c.jsonl'):
    try:
        d = json.loads(l); print('%-42s W %5s %8.3f M moves/s %7.2f ms C1 %.4f N %.1f' % (d['lib'][-40:], d.get('walkers'), d['moves_per_s']/1e6, d['ms_per_step'], d.get('frac_c1', 0), d['loading'][1]))
    except Exception as e: print('bad', l[:80])
PY
