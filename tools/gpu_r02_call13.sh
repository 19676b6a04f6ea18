#!/bin/bash
V=${1:-r02_c13}
O=gpurun_out
mkdir -p $O
nproc > $O/nproc_$V.txt
timeout 1500 python -m pytest tests -q -m gpu --durations=8 -s > $O/pytest_gpu_$V.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_$V.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$V.log 2>&1; echo "rc=$?" >> $O/smoke_$V.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-step > $O/bench_$V.json 2> $O/bench_${V}_kernel_breakdown.txt; echo "rc=$?" >> $O/bench_${V}_kernel_breakdown.txt
