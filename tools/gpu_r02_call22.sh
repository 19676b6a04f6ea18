#!/bin/bash
V=${1:-r02_c22}
O=gpurun_out
timeout 500 python -m pytest tests/test_gemm_gpu.py tests/test_unet_gpu.py tests/test_clip_gpu.py tests/test_zz_batch_invariance_gpu.py -q -x -k "not full_size" > $O/pytest_tmastore_$V.log 2>&1; echo "rc=$?" >> $O/pytest_tmastore_$V.log
for i in 1 2; do
  DFB_GEMM_TMA_STORE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --profile-step > $O/ab_tmastore_off_$i.json 2> $O/ab_tmastore_off_${i}_breakdown.txt
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --profile-step > $O/ab_tmastore_on_$i.json 2> $O/ab_tmastore_on_${i}_breakdown.txt
done
