#!/bin/bash
# Headline-path capture (bf16 tcgen05 kernels, minibatch 32768), run on the GPU box through gpurun:
#   gpurun --timeout 900 -- 'bash profiles/capture_bf16.sh r01p'
# GPU tests, the default bench line, then the two ncu passes of B200_PROFILING.md — each only after the same command
# has exited 0 without ncu.  Raw output lands in gpurun_out/; profiles/summarize.py turns it into the tracked files.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1
echo tests_rc=$?
tail -3 $OUT/${TAG}_tests.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo bench_rc=$?
SHORT="python bench.py --steps 3 --warmup 3 --no-kernels --no-cpu --no-variants"
BENCH="python bench.py --steps 1 --warmup 3 --epochs 1 --no-kernels --no-cpu"
$BENCH > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_launches.log 2>&1
echo launches_rc=$?
K="python profiles/kernels.py update --precision bf16 --batch 32768"
$K > $OUT/${TAG}_k_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:tc_|adam_cast' -s 7 -c 7 -f -o $OUT/${TAG}_tc $K > $OUT/${TAG}_k_ncu.log 2>&1
echo tc_rc=$?
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1
echo smoke_rc=$?
