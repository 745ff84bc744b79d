#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=r03l
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --tb=short -x -k "mixture_triclinic or skewed or dipole_triclinic or sweep_mixture" --durations=5 > gpurun_out/pytest_$TAG.log 2>&1
tail -9 gpurun_out/pytest_$TAG.log
timeout 200 python tools/mixture_probe.py 2368 32
timeout 200 python tools/mixture_probe.py 2368 32
