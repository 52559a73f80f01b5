#!/bin/bash
# Run on the GPU box through gpurun: plain run first (must exit 0), then the ncu passes (B200_PROFILING.md).
# usage: bash profiles/capture.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 3 --epochs 1 --no-kernels --no-cpu"
$BENCH > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1400 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_launches.log 2>&1
echo launches_rc=$?
for K in gae gather adam update; do
  case $K in
    gae) PAT=gae_scan_kernel;; gather) PAT=gather_minibatch_kernel;; adam) PAT=adam_kernel;; update) PAT=${UPDATE_PAT:-gemm_group_kernel};;
  esac
  SKIP=1; CNT=2
  if [ $K = update ]; then SKIP=${UPDATE_SKIP:-20}; CNT=${UPDATE_CNT:-9}; fi
  python profiles/kernels.py $K ${KARGS:-} > $OUT/${TAG}_${K}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$PAT -s $SKIP -c $CNT -f -o $OUT/${TAG}_${K} python profiles/kernels.py $K ${KARGS:-} > $OUT/${TAG}_${K}_ncu.log 2>&1
  echo ${K}_rc=$?
done
ls -la $OUT | tail -20
