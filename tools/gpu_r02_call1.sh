#!/bin/bash
# Round 2, call 1: locate the shared-prefix bit mismatch (diag + sanitizers), whole GPU suite without -x, pending
# verifications of round 1, bench baseline on this box, ping-pong attention experiment (last: a faulty kernel poisons its process).
V=${1:-r02_c1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu_$V.txt 2>&1
timeout 400 python tools/shared_prefix_diag.py 4 > $O/shared_prefix_diag_$V.log 2>&1; echo "rc=$?" >> $O/shared_prefix_diag_$V.log
timeout 900 python -m pytest tests -q -m gpu --durations=12 > $O/pytest_gpu_$V.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_$V.log
timeout 120 python tools/batch_invariance_diag.py > $O/batch_invariance_$V.log 2>&1
timeout 120 python tools/plms_chunk_diag.py > $O/plms_chunk_$V.log 2>&1
T='tests/test_unet_gpu.py::test_shared_cfg_prefix_is_bitwise_the_full_computation'
for tool in memcheck racecheck initcheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 40 python -m pytest "$T" -x -q -k "scales7" > $O/sanitizer_${tool}_$V.log 2>&1
  echo "rc=$?" >> $O/sanitizer_${tool}_$V.log
done
timeout 600 python bench.py --profile-step > $O/bench_$V.json 2> $O/bench_${V}_kernel_breakdown.txt; echo "rc=$?" >> $O/bench_${V}_kernel_breakdown.txt
DFB_TEST_PP=1 timeout 200 python -m pytest tests/test_attn_gpu.py -q -k ping_pong > $O/pytest_pp_$V.log 2>&1; echo "rc=$?" >> $O/pytest_pp_$V.log
timeout 200 python tools/attn_pp_experiment.py > $O/attn_pp_$V.log 2>&1; echo "rc=$?" >> $O/attn_pp_$V.log
ls -la $O | tail -30
