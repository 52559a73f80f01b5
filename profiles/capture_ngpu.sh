#!/bin/bash
# N-GPU bench line, run through `gpurun --gpus N`: one rank per GPU, launched the way the driver launches it.
#   gpurun --gpus 4 --timeout 400 -- 'bash profiles/capture_ngpu.sh 4 r01s'
set -u
N=${1:-2}
TAG=${2:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo_${N}gpu.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 \
  bench.py --gpus $N --steps 3 --warmup 3 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
echo bench_${N}gpu_rc=$?
tail -c 400 $OUT/${TAG}_bench_${N}gpu.json
