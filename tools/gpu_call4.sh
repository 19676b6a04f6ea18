#!/bin/bash
# GPU call 4: fused upsample phases — kernel tests, UNet parity, A/B bench (interleaved).
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -k "upsample" > $O/pytest_call4_gemm.log 2>&1; echo "rc=$?" >> $O/pytest_call4_gemm.log
timeout 400 python -m pytest tests/test_unet_gpu.py tests/test_streaming_gpu.py -q --maxfail=20 > $O/pytest_call4_unet.log 2>&1; echo "rc=$?" >> $O/pytest_call4_unet.log
for i in 1 2; do
  DFB_UPSAMPLE_PHASES=0 timeout 300 python bench.py --steps 20 --no-e2e --no-cpu-baseline > $O/ab4_literal_$i.json 2>> $O/ab4.err
  timeout 300 python bench.py --steps 20 --no-e2e --no-cpu-baseline --profile-step > $O/ab4_phases_$i.json 2>> $O/ab4_phases_$i.txt
done
ls -la $O | tail -12
