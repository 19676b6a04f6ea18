#!/bin/bash
V=${1:-r02_c4}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_attn_gpu.py -q -x > $O/pytest_attn_$V.log 2>&1; echo "rc=$?" >> $O/pytest_attn_$V.log
timeout 200 python tools/attn_sa_experiment.py > $O/attn_sa_$V.log 2>&1; echo "rc=$?" >> $O/attn_sa_$V.log
timeout 400 python -m pytest tests/test_unet_gpu.py -q -x -k "forward_matches or shared_cfg or weight_changes or fitb" > $O/pytest_unet_$V.log 2>&1; echo "rc=$?" >> $O/pytest_unet_$V.log
