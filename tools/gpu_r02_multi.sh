#!/bin/bash
# Multi-GPU evidence (run with gpurun --gpus N): N = 2: NCCL two-rank bitwise test + C2/C4 lines; N = 4: C4; N = 8: C2, C4 (with
# the gathered e2e leg), C5 sweep.  Usage: bash tools/gpu_r02_multi.sh N tag
N=$1; V=${2:-r02}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L > $O/gpus_${V}_n$N.txt 2>&1
if [ "$N" = "2" ]; then
  DFB_TEST_NCCL=1 timeout 600 python -m pytest tests/test_sharded_gpu.py -q -s > $O/pytest_sharded_nccl_$V.log 2>&1; echo "rc=$?" >> $O/pytest_sharded_nccl_$V.log
fi
timeout 600 $TR --master-port 29501 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_${V}_C2_n$N.json 2> $O/bench_${V}_C2_n$N.err; echo "rc=$?" >> $O/bench_${V}_C2_n$N.err
if [ "$N" = "8" ]; then
  timeout 900 $TR --master-port 29502 bench.py --gpus $N --config C4 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${V}_C4_n$N.json 2> $O/bench_${V}_C4_n$N.err; echo "rc=$?" >> $O/bench_${V}_C4_n$N.err
  timeout 900 $TR --master-port 29503 tools/sweep_c5.py --out $O/c5_sweep_${V}_n$N.jsonl > $O/c5_sweep_${V}_n$N.log 2>&1; echo "rc=$?" >> $O/c5_sweep_${V}_n$N.log
else
  timeout 900 $TR --master-port 29502 bench.py --gpus $N --config C4 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_${V}_C4_n$N.json 2> $O/bench_${V}_C4_n$N.err; echo "rc=$?" >> $O/bench_${V}_C4_n$N.err
fi
ls -la $O | tail -8
