#!/bin/bash
# final validation of a round: full GPU test suite, the full bench line, a capture of the triclinic sweep
set -u
mkdir -p gpurun_out
TAG=${1:-final}
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --tb=short > gpurun_out/pytest_$TAG.log 2>&1
tail -4 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/prof_mixture_$TAG -f python tools/mixture_probe.py 2368 16 > gpurun_out/ncu_mixture_$TAG.log 2>&1
tail -2 gpurun_out/ncu_mixture_$TAG.log
