#!/bin/bash
V=${1:-r02_c20}
O=gpurun_out
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_unet_gpu.py -q -x -k "not full_size and not sd2" > $O/pytest_direct_$V.log 2>&1; echo "rc=$?" >> $O/pytest_direct_$V.log
for i in 1 2; do
  for lib in base geglu_only both; do
    L=$PWD/difashion_b200/libdfb200_$lib.so; [ "$lib" = "both" ] && L=$PWD/difashion_b200/libdfb200.so
    DFB200_LIB=$L timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --profile-step > $O/ab_direct_${lib}_$i.json 2> $O/ab_direct_${lib}_${i}_breakdown.txt
  done
done
