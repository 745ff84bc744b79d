#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=r03k
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --tb=short -x -k "triclinic or mixture or skewed" > gpurun_out/pytest_$TAG.log 2>&1
tail -3 gpurun_out/pytest_$TAG.log
V=$PWD/maniac-mc.github.io_b200/variants
timeout 200 python tools/mixture_probe.py 2368 32
timeout 200 python tools/mixture_probe.py 2368 32 5
export MANIAC_GPU_LIB=$V/libmaniac_gpu_base.so
timeout 200 python tools/mixture_probe.py 2368 32
unset MANIAC_GPU_LIB
timeout 300 python bench.py --quick --steps 6 --warmup 3 --walkers 4736
