#!/bin/bash
# gpurun --timeout 600 -- 'bash tools/dev/run_gemm_ncu.sh'   -> gpurun_out/gemm/tc_full.ncu-rep (+ csv pages)
mkdir -p gpurun_out/gemm
export LD_LIBRARY_PATH=$PWD/hotrack_b200:$LD_LIBRARY_PATH
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 11 --launch-count 1 \
   -o gpurun_out/gemm/tc_fwd -f ./tools/dev/gemm_tc_check prof > gpurun_out/gemm/ncu_fwd.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 23 --launch-count 1 \
   -o gpurun_out/gemm/tc_dgrad -f ./tools/dev/gemm_tc_check prof > gpurun_out/gemm/ncu_dgrad.log 2>&1
ls -la gpurun_out/gemm; tail -3 gpurun_out/gemm/ncu_fwd.log
