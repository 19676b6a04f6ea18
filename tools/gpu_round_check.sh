#!/bin/bash
# One GPU-box call that refreshes the round's evidence: new parity tests first, the whole GPU suite, smoke(), the bench line with
# its per-kernel breakdown, the reference arm, the ncu launch list of one step, and a short outfit-count sweep (BASELINE configs[4]).
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round_check.sh v15'
V=${1:-vNN}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu_$V.txt 2>&1
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_clip_gpu.py -q -k "sd2 or gelu" -s > $O/pytest_new_$V.log 2>&1; echo "rc=$?" >> $O/pytest_new_$V.log
timeout 1000 python -m pytest tests -q -m gpu --maxfail=10 --durations=12 > $O/pytest_gpu_$V.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_$V.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$V.log 2>&1; echo "rc=$?" >> $O/smoke_$V.log
timeout 600 python bench.py --profile-step > $O/bench_$V.json 2> $O/bench_${V}_kernel_breakdown.txt; echo "rc=$?" >> $O/bench_${V}_kernel_breakdown.txt
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$V.json 2> $O/bench_ref_$V.err
timeout 600 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --csv --log-file $O/ncu_launches_step_$V.csv python tools/step_traffic.py > $O/step_traffic_$V.log 2>&1
python tools/step_traffic.py --summarise $O/ncu_launches_step_$V.csv $O/step_traffic_$V.json >> $O/step_traffic_$V.log 2>&1
for n in ${SWEEP:-1 4 64}; do
  timeout 300 python bench.py --outfits $n --steps 10 --no-e2e --no-cpu-baseline > $O/bench_${V}_sweep_${n}outfits.json 2>> $O/sweep_$V.err
done
ls -la $O
