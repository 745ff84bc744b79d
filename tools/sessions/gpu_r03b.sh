#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=r03b
OUT=gpurun_out/variants_$TAG.jsonl; : > $OUT
V=$PWD/maniac-mc.github.io_b200/variants
run() { # lib walkers
  if [ "$1" = default ]; then unset MANIAC_GPU_LIB; else export MANIAC_GPU_LIB=$V/libmaniac_gpu_$1.so; fi
  timeout 300 python bench.py --quick --steps 6 --warmup 3 --walkers $2 >> $OUT 2>> gpurun_out/variants_$TAG.err
}
for rep in 1 2; do
  run base 4736; run default 4736; run ag8 4736; run ag2 4736
done
unset MANIAC_GPU_LIB
python - <<'PY'
import json
for l in open('gpurun_out/variants_r03b.jsonl'):
    try:
        d = json.loads(l); print('%-42s %8.3f M moves/s %7.2f ms C1 %.4f N %.1f sm %s' % (d['lib'][-40:], d['moves_per_s']/1e6, d['ms_per_step'], d.get('frac_c1', 0), d['loading'][1], d['clocks'].get('sm_mhz')))
    except Exception as e: print('bad', l[:80])
PY
M=gpurun_out/mixture_$TAG.log; : > $M
export MANIAC_GPU_LIB=$V/libmaniac_gpu_ag8.so
timeout 200 python tools/mixture_probe.py 2368 32 >> $M 2>&1
unset MANIAC_GPU_LIB
timeout 200 python tools/mixture_probe.py 2368 32 >> $M 2>&1
timeout 200 python tools/mixture_probe.py 2368 32 0 >> $M 2>&1
cat $M
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/prof_mixture_$TAG -f python tools/mixture_probe.py 2368 16 > gpurun_out/ncu_mixture_$TAG.log 2>&1
tail -3 gpurun_out/ncu_mixture_$TAG.log
