#!/bin/bash
# Round-2 capture of the fp32-tolerance path on tcgen05 (minibatch 32768: split3_kernel + tc_persist_kernel in three-term mode +
# heads_fused_kernel + adam_kernel), run on the GPU box through gpurun:  gpurun --timeout 900 -- 'bash profiles/capture_r02_fp32.sh'
# The ncu pass only after the same command has exited 0 without ncu (B200_PROFILING.md); one ncu pass per call.
set -u
OUT=gpurun_out
mkdir -p $OUT
K="python profiles/kernels.py update --precision fp32 --batch 32768"
$K > $OUT/r02_fp32_k_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:split3|tc_gemm_kernel|tc_persist|heads_fused|adam' -s 10 -c 10 -f -o $OUT/r02_fp32_tc $K > $OUT/r02_fp32_k_ncu.log 2>&1
echo fp32_tc_rc=$?
