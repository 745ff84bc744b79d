#!/bin/bash
# Kernel-variant shoot-out on one GPU box: tools/gpu_variants.sh TAG variant1 variant2 ...
# (libraries built by tools/build_variant.sh; "default" = the in-tree libmaniac_gpu.so)
set -u
mkdir -p gpurun_out
TAG=$1; shift
OUT=gpurun_out/variants_$TAG.jsonl
: > $OUT
for v in "$@"; do
  if [ "$v" = default ]; then unset MANIAC_GPU_LIB; else export MANIAC_GPU_LIB=$PWD/maniac-mc.github.io_b200/variants/libmaniac_gpu_$v.so; fi
  for rep in 1 2; do
    timeout 300 python bench.py --quick ${QUICK_ARGS:---steps 6 --warmup 3} >> $OUT 2>> gpurun_out/variants_$TAG.err
  done
done
unset MANIAC_GPU_LIB
cat $OUT | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print('%-50s %8.3f M moves/s  %7.2f ms  C1 %.4f  N %.1f  sm %s' % (d['lib'][-40:], d['moves_per_s']/1e6, d['ms_per_step'], d.get('frac_c1', 0), d['loading'][1], d['clocks'].get('sm_mhz')))
    except Exception as e: print('bad line', l[:100])
"
