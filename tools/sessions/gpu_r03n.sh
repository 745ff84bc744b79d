#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=r03n
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --tb=short -x -k "mixture_triclinic or skewed or dipole_triclinic or sweep_mixture or (large_triclinic and hcache)" --durations=3 > gpurun_out/pytest_$TAG.log 2>&1
tail -6 gpurun_out/pytest_$TAG.log
timeout 200 python tools/mixture_probe.py 2368 32
timeout 200 python tools/mixture_probe.py 2368 32
