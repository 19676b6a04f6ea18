#!/bin/bash
V=${1:-r02_c19}
O=gpurun_out
timeout 60 ./tools/ubench/gelu_tanh_accuracy > $O/gelu_tanh_accuracy_$V.log 2>&1
for lib in sigmoid tanh; do
  DFB200_LIB=$PWD/difashion_b200/libdfb200_$lib.so timeout 300 python -m pytest tests/test_unet_gpu.py tests/test_gemm_gpu.py -q -s -k "forward_matches_oracle or geglu" > $O/pytest_geglu_${lib}_$V.log 2>&1; echo "rc=$?" >> $O/pytest_geglu_${lib}_$V.log
done
for i in 1 2; do
  for lib in sigmoid tanh; do
    DFB200_LIB=$PWD/difashion_b200/libdfb200_$lib.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --profile-step > $O/ab_geglu_${lib}_$i.json 2> $O/ab_geglu_${lib}_${i}_breakdown.txt
  done
  DFB_ATTN_R01=1 DFB200_LIB=$PWD/difashion_b200/libdfb200_sigmoid.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/ab_attn_r01_$i.json 2>> $O/ab_$V.err
done
