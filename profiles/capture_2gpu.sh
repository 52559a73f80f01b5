#!/bin/bash
# Two-GPU check, run through `gpurun --gpus 2`: the data-parallel parity tests (NCCL and peer-memory exchange) and
# the bench line at N=2 (one rank per GPU, launched the way the driver launches it).
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_distributed.py -m gpu -q > $OUT/${TAG}_dist_tests.log 2>&1
echo dist_tests_rc=$?
tail -3 $OUT/${TAG}_dist_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/${TAG}_bench_2gpu.json 2> $OUT/${TAG}_bench_2gpu.err
echo bench_2gpu_rc=$?
tail -c 600 $OUT/${TAG}_bench_2gpu.json
