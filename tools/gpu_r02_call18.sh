#!/bin/bash
V=${1:-r02_c18}
O=gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x -s -k "layernorm_folded or upsample" > $O/pytest_lnfold_$V.log 2>&1; echo "rc=$?" >> $O/pytest_lnfold_$V.log
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_zz_batch_invariance_gpu.py -q -x -s -k "not full_size" > $O/pytest_unet_$V.log 2>&1; echo "rc=$?" >> $O/pytest_unet_$V.log
for i in 1 2; do
  DFB_LN_FOLD=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/ab_lnfold_off_$i.json 2>> $O/ab_lnfold_$V.err
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --profile-step > $O/ab_lnfold_on_$i.json 2> $O/ab_lnfold_on_${i}_breakdown.txt
done
