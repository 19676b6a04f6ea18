#!/bin/bash
# GPU call 2 of the session: new bitwise tests of the shared CFG prefix, A/B bench (interleaved), ncu full of the wide-tile conv.
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_generation_gpu.py tests/test_attn_gpu.py -q --maxfail=20 > $O/pytest_call2.log 2>&1; echo "rc=$?" >> $O/pytest_call2.log
for i in 1 2; do
  timeout 300 python bench.py --steps 20 --no-e2e --no-cpu-baseline --no-share-prefix > $O/ab_noshare_$i.json 2>> $O/ab.err
  timeout 300 python bench.py --steps 20 --no-e2e --no-cpu-baseline > $O/ab_share_$i.json 2>> $O/ab.err
done
WHICH=conv320wide ROWS=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 --launch-skip 2 --launch-count 2 \
  -f -o $O/ncu_full_v15_wide python tools/ncu_targets.py > $O/ncu_full_v15_wide.log 2>&1
ncu -i $O/ncu_full_v15_wide.ncu-rep --page raw --csv > $O/ncu_full_v15_wide_raw.csv 2>> $O/ncu_full_v15_wide.log
ls -la $O
