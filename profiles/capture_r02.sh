#!/bin/bash
# Round-2 capture of the headline path (bf16, minibatch 32768: tc_chain_kernel + tc_wgrad2_kernel + adam_cast_kernel), run on
# the GPU box through gpurun:   gpurun --timeout 900 -- 'bash profiles/capture_r02.sh r02'
# Each ncu pass only after the same command has exited 0 without ncu (B200_PROFILING.md).
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 3 --epochs 1 --no-kernels --no-cpu --no-variants --no-e2e"
$BENCH > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_launches.log 2>&1
echo launches_rc=$?
K="python profiles/kernels.py update --precision bf16 --batch 32768"
$K > $OUT/${TAG}_k_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:tc_chain|tc_wgrad2|adam_cast' -s 3 -c 3 -f -o $OUT/${TAG}_tc $K > $OUT/${TAG}_k_ncu.log 2>&1
echo tc_rc=$?
