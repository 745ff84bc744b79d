#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 200 --tb=short -x -k "mixture_triclinic or skewed or dipole_triclinic or automatic_sweep_shapes" > gpurun_out/pytest_r03q.log 2>&1
tail -3 gpurun_out/pytest_r03q.log
timeout 100 python tools/mixture_probe.py 2368 32
