#!/bin/bash
# First GPU contact: staged so that one failing stage does not hide the others.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1; free -g >> gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" | tee -a gpurun_out/summary.txt; tail -5 gpurun_out/$name.log; }
TMO=300 run t_dev_basic python -m pytest tests/test_gpu_device.py -q -x -k "synth or decode or transpose or argmax or gemv" -p no:cacheprovider
TMO=300 run t_readblock python -m pytest tests/test_gpu_parity.py -q -k "readblock" -p no:cacheprovider
TMO=300 run t_mmt python -m pytest tests/test_gpu_parity.py -q -k "mmt" -p no:cacheprovider
TMO=300 run t_scan python -m pytest tests/test_gpu_parity.py -q -k "scan or reduced" -p no:cacheprovider
TMO=600 run t_rest python -m pytest tests/test_gpu_parity.py -q -k "not mmt and not scan and not reduced and not readblock" -p no:cacheprovider
TMO=300 run smoke python __graft_entry__.py smoke
TMO=600 run t_c2 python -m pytest tests/test_gpu_device.py -q -k "config2" -p no:cacheprovider
TMO=600 run bench_c2 python bench.py --workload c2 --steps 2 --warmup 1
cat gpurun_out/summary.txt
