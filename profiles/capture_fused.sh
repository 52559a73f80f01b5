#!/bin/bash
# A/B of the output-layer dgrad inside the fused-loss epilogue (B200PPO_FUSE_OUT_DGRAD=1) on the GPU box:
#   gpurun --timeout 600 -- 'bash profiles/capture_fused.sh r01q'
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
FUSED=test_bf16_output_dgrad_in_the_loss_epilogue_matches_the_separate_launch
timeout 300 python -m pytest tests/test_update_gpu.py -m gpu -q -k "$FUSED" > $OUT/${TAG}_tests_fused.log 2>&1
echo tests_fused_rc=$?
tail -5 $OUT/${TAG}_tests_fused.log
SHORT="python bench.py --steps 3 --warmup 3 --no-kernels --no-cpu --no-variants"
timeout 300 $SHORT > $OUT/${TAG}_bench_sep.json 2> $OUT/${TAG}_bench.err
echo bench_sep_rc=$?
B200PPO_FUSE_OUT_DGRAD=1 timeout 300 $SHORT > $OUT/${TAG}_bench_fused.json 2>> $OUT/${TAG}_bench.err
echo bench_fused_rc=$?
for COST in 3,1 2,1 6,1; do
  B200PPO_FUSE_OUT_DGRAD=1 B200PPO_FUSE_COST=$COST timeout 300 $SHORT --no-e2e > $OUT/${TAG}_bench_fused_cost_${COST/,/_}.json 2>> $OUT/${TAG}_bench.err
  echo bench_fused_cost_${COST}_rc=$?
done
B200PPO_FUSE_OUT_DGRAD=1 timeout 300 python -m pytest tests/test_update_gpu.py tests/test_distributed.py -m gpu -q -k "bf16" > $OUT/${TAG}_tests_bf16_fused_env.log 2>&1
echo tests_bf16_env_rc=$?
tail -3 $OUT/${TAG}_tests_bf16_fused_env.log
BENCH="python bench.py --steps 1 --warmup 3 --epochs 1 --no-kernels --no-cpu --no-variants"
B200PPO_FUSE_OUT_DGRAD=1 $BENCH > $OUT/${TAG}_plain_fused.log 2>&1 &&
B200PPO_FUSE_OUT_DGRAD=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_fused.csv $BENCH > $OUT/${TAG}_ncu_launches_fused.log 2>&1
echo launches_fused_rc=$?
