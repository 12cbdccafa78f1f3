#!/bin/bash
# gpurun --timeout 300 -- 'bash tools/dev/run_gemm_check.sh [time]'
mkdir -p gpurun_out/gemm
export LD_LIBRARY_PATH=$PWD/hotrack_b200:$LD_LIBRARY_PATH
PN2_WGRAD_IMPL=tc PN2_GEMM_IMPL=tc timeout 90 ./tools/dev/gemm_tc_check $1 > gpurun_out/gemm/tc.log 2>&1; echo "tc rc=$?" >> gpurun_out/gemm/tc.log
PN2_GEMM_IMPL=mma timeout 90 ./tools/dev/gemm_tc_check $1 > gpurun_out/gemm/mma.log 2>&1; echo "mma rc=$?" >> gpurun_out/gemm/mma.log
cat gpurun_out/gemm/tc.log; echo ------; cat gpurun_out/gemm/mma.log
