#!/bin/bash
# Round 2, call 3: pipe micro-benchmark for the attention softmax loop, the whole GPU suite after the two rounding fixes
# (blend FMA contraction, generic-epilogue row bias order), the diagnostics again, smoke, new bench.py (default + a C3 line).
V=${1:-r02_c3}
O=gpurun_out
mkdir -p $O
timeout 120 ./tools/ubench/ubench_softmax_pipes > $O/ubench_softmax_pipes_$V.log 2>&1; echo "rc=$?" >> $O/ubench_softmax_pipes_$V.log
timeout 900 python -m pytest tests -q -m gpu --durations=8 > $O/pytest_gpu_$V.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_$V.log
timeout 300 python tools/shared_prefix_diag.py 2 > $O/shared_prefix_diag_$V.log 2>&1; echo "rc=$?" >> $O/shared_prefix_diag_$V.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$V.log 2>&1; echo "rc=$?" >> $O/smoke_$V.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-step > $O/bench_$V.json 2> $O/bench_${V}_kernel_breakdown.txt; echo "rc=$?" >> $O/bench_${V}_kernel_breakdown.txt
timeout 300 python bench.py --config C3 --steps 10 --no-cpu-baseline > $O/bench_${V}_C3.json 2> $O/bench_${V}_C3.err; echo "rc=$?" >> $O/bench_${V}_C3.err
ls -la $O | tail -12
