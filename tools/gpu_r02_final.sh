#!/bin/bash
# Last GPU call of the round: exactly what the driver runs, on HEAD.
V=${1:-r02_final}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu_$V.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu_$V.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_$V.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$V.log 2>&1; echo "rc=$?" >> $O/smoke_$V.log
timeout 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_ref_$V.json 2> $O/bench_ref_$V.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_$V.json 2> $O/bench_$V.err; echo "rc=$?" >> $O/bench_$V.err
