#!/bin/bash
V=${1:-r02_c9}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu --durations=5 > $O/pytest_gpu_$V.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_$V.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$V.log 2>&1; echo "rc=$?" >> $O/smoke_$V.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-step > $O/bench_$V.json 2> $O/bench_${V}_kernel_breakdown.txt; echo "rc=$?" >> $O/bench_${V}_kernel_breakdown.txt
timeout 300 python bench.py --config C3 --steps 10 --no-cpu-baseline > $O/bench_${V}_C3.json 2> $O/bench_${V}_C3.err; echo "rc=$?" >> $O/bench_${V}_C3.err
