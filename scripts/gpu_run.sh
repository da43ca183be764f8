#!/bin/bash
# One parameterised launcher for gpurun (replaces the per-call gpu_NN.sh files of round 1).
#   scripts/gpu_run.sh tests            -> pytest -m gpu
#   scripts/gpu_run.sh bench [args...]  -> python bench.py args (N = 1)
#   scripts/gpu_run.sh benchN N [args]  -> torchrun bench.py --gpus N args
# Output goes to gpurun_out/<tag>.{log,json}; TAG env names the files.
set -u
mkdir -p gpurun_out
TAG=${TAG:-run}
what=$1; shift
case "$what" in
  tests) timeout ${T:-1500} python -m pytest tests -m gpu -x -q "$@" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log ;;
  bench) timeout ${T:-900} python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench.err; head -c 3000 gpurun_out/${TAG}_bench.json ;;
  benchN) N=$1; shift; timeout ${T:-1200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "benchN rc=$?"; tail -c 2500 gpurun_out/${TAG}_bench_n$N.err; head -c 3000 gpurun_out/${TAG}_bench_n$N.json ;;
  *) echo "unknown $what"; exit 2 ;;
esac
