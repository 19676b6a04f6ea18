#!/bin/bash
# Round-2 evidence refresh on ONE GPU box: whole GPU suite, smoke, bench line + per-kernel breakdown, reference arm, ncu launch list +
# DRAM traffic of one step, sanitizer pass over the attention tests and smoke, C4 strong-scaling N=1 point, C3 line.
V=${1:-r02_final}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu_$V.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --durations=8 -s > $O/pytest_gpu_$V.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_$V.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$V.log 2>&1; echo "rc=$?" >> $O/smoke_$V.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-step > $O/bench_$V.json 2> $O/bench_${V}_kernel_breakdown.txt; echo "rc=$?" >> $O/bench_${V}_kernel_breakdown.txt
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref_$V.json 2> $O/bench_ref_$V.err
timeout 600 ncu --profile-from-start off --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --csv --log-file $O/ncu_launches_step_$V.csv python tools/step_traffic.py > $O/step_traffic_$V.log 2>&1
python tools/step_traffic.py --summarise $O/ncu_launches_step_$V.csv $O/step_traffic_$V.json >> $O/step_traffic_$V.log 2>&1
timeout 300 python bench.py --config C3 --steps 10 --no-cpu-baseline > $O/bench_${V}_C3.json 2> $O/bench_${V}_C3.err
timeout 300 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_${V}_C4_n1.json 2> $O/bench_${V}_C4_n1.err
for tool in memcheck initcheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_attn_gpu.py -x -q -k "ones_column or eight_warp or deterministic or redo" > $O/sanitizer_${tool}_attn_$V.log 2>&1; echo "rc=$?" >> $O/sanitizer_${tool}_attn_$V.log
done
# timing robustness of every long-sequence attention variant under the sanitizer's timing (racecheck found the tile-split epilogue race)
REPS=3 timeout 600 compute-sanitizer --tool racecheck python tools/attn_race_probe.py > $O/race_probe_racecheck_$V.log 2>&1; echo "rc=$?" >> $O/race_probe_racecheck_$V.log
DFB_SKIP_SLOW=1 timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_unet_gpu.py -x -q -k "tiny-4-77 or shared_cfg or fitb" > $O/sanitizer_memcheck_unet_$V.log 2>&1; echo "rc=$?" >> $O/sanitizer_memcheck_unet_$V.log
ls -la $O | tail -16
