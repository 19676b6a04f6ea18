#!/bin/bash
# First GPU call of the next round: (1) the items of session 5 that never ran on hardware — the batch-invariant GroupNorm
# statistics decomposition + tests/test_zz_batch_invariance_gpu.py, PLMS over several row chunks; (2) the whole GPU suite,
# smoke(), the bench line with its breakdown, the reference arm and the ncu launch list (tools/gpu_round_check.sh).
# Usage: gpurun --timeout 1200 -- 'bash tools/gpu_next_round.sh r02_v1'
V=${1:-r02_v1}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_zz_batch_invariance_gpu.py tests/test_unet_gpu.py -q -k "batch or plms or chunk" -s > $O/pytest_pending_$V.log 2>&1; echo "rc=$?" >> $O/pytest_pending_$V.log
timeout 120 python tools/batch_invariance_diag.py > $O/batch_invariance_$V.log 2>&1
timeout 120 python tools/plms_chunk_diag.py > $O/plms_chunk_$V.log 2>&1
bash tools/gpu_round_check.sh $V
# last, because a faulty kernel traps and poisons its process: the ping-pong attention experiment (dbg bit12, never run)
DFB_TEST_PP=1 timeout 200 python -m pytest tests/test_attn_gpu.py -q -k ping_pong > $O/pytest_pp_$V.log 2>&1; echo "rc=$?" >> $O/pytest_pp_$V.log
timeout 200 python tools/attn_pp_experiment.py > $O/attn_pp_$V.log 2>&1; echo "rc=$?" >> $O/attn_pp_$V.log
