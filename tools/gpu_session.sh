#!/bin/bash
# One GPU-box session.  usage: tools/gpu_session.sh TAG [test] [bench] [prof] [extra...]
# Outputs land in gpurun_out/ (copied back by gpurun).
set -u
mkdir -p gpurun_out
TAG=${1:-r01}; shift
BENCH_ARGS=${BENCH_ARGS:-"--steps 8 --warmup 3"}
BENCH_SMALL="python bench.py --steps 2 --warmup 1 --walkers 2368 --inner ${PROF_INNER:-64} --widom 200000 --e2e-walkers 256 --e2e-steps 3 --no-cpu --mixture-walkers 0 --grid 0 --equil 12 --strong-walkers 0"
for what in "$@"; do
case $what in
test)
  timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --tb=short ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_$TAG.log 2>&1
  tail -40 gpurun_out/pytest_$TAG.log ;;
bench)
  timeout 900 python bench.py $BENCH_ARGS > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
  tail -c 4000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err ;;
prof)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH_SMALL > gpurun_out/ncu_launch_$TAG.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/prof_sweep_$TAG -f $BENCH_SMALL > gpurun_out/ncu_full_$TAG.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_widom -s 1 -c 1 -o gpurun_out/prof_widom_$TAG -f $BENCH_SMALL >> gpurun_out/ncu_full_$TAG.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 1 -c 1 -o gpurun_out/prof_mixture_$TAG -f python tools/mixture_probe.py 2368 16 >> gpurun_out/ncu_full_$TAG.log 2>&1
  ls -la gpurun_out/ ;;
*)
  echo "unknown step $what" ;;
esac
done
