#!/bin/bash
# session r03a: parity of the new k-space / triclinic gate, then variant shoot-out (ortho + mixture)
set -u
mkdir -p gpurun_out
TAG=r03a
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --tb=short -x > gpurun_out/pytest_$TAG.log 2>&1
tail -5 gpurun_out/pytest_$TAG.log
OUT=gpurun_out/variants_$TAG.jsonl; : > $OUT
V=$PWD/maniac-mc.github.io_b200/variants
run() { # lib walkers
  if [ "$1" = default ]; then unset MANIAC_GPU_LIB; else export MANIAC_GPU_LIB=$V/libmaniac_gpu_$1.so; fi
  timeout 300 python bench.py --quick --steps 6 --warmup 3 --walkers $2 >> $OUT 2>> gpurun_out/variants_$TAG.err
}
for rep in 1 2; do
  run base 4736; run default 4736; run swp 4736; run w384 5328; run w384swp 5328
done
unset MANIAC_GPU_LIB
python - <<'PY'
import json
for l in open('gpurun_out/variants_r03a.jsonl'):
    try:
        d = json.loads(l); print('%-42s %8.3f M moves/s %7.2f ms C1 %.4f N %.1f sm %s' % (d['lib'][-40:], d['moves_per_s']/1e6, d['ms_per_step'], d.get('frac_c1', 0), d['loading'][1], d['clocks'].get('sm_mhz')))
    except Exception as e: print('bad', l[:80])
PY
M=gpurun_out/mixture_$TAG.log; : > $M
for lib in base default; do
  if [ "$lib" = default ]; then unset MANIAC_GPU_LIB; else export MANIAC_GPU_LIB=$V/libmaniac_gpu_$lib.so; fi
  timeout 200 python tools/mixture_probe.py 2368 32 >> $M 2>&1
done
unset MANIAC_GPU_LIB
for ps in 1 3 7 11; do timeout 200 python tools/mixture_probe.py 2368 32 $ps >> $M 2>&1; done
cat $M
