#!/bin/bash
# gpurun --timeout 600 -- 'bash tools/dev/run_gemm_ncu.sh <mode> <kernel regex> <skip>'   -> gpurun_out/gemm/<mode>.ncu-rep
mkdir -p gpurun_out/gemm
export LD_LIBRARY_PATH=$PWD/hotrack_b200:$LD_LIBRARY_PATH
PN2_WGRAD_IMPL=tc timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip $3 --launch-count 1 \
   -o gpurun_out/gemm/$1 -f ./tools/dev/gemm_tc_check $1 > gpurun_out/gemm/ncu_$1.log 2>&1
ls -la gpurun_out/gemm; tail -3 gpurun_out/gemm/ncu_$1.log
