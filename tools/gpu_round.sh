#!/bin/bash
# One gpurun call: [GPU tests,] both bench arms, ncu launch list, ncu --set full of the hot kernels.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag> [tests] [ref] [full]'
tag=${1:-run}; shift
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
for what in "$@"; do case $what in
tests) timeout 900 python -m pytest tests -m gpu -x -q > $out/tests.log 2>&1; echo "tests rc=$?" >> $out/tests.log; tail -3 $out/tests.log;;
ref) timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $out/bench_ref.json 2> $out/bench_ref.err; cut -c1-300 $out/bench_ref.json;;
bench) timeout 400 python bench.py > $out/bench_ours.json 2> $out/bench_ours.err; cut -c1-300 $out/bench_ours.json;;
launches) timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${LAUNCH_SKIP:-480} --launch-count ${LAUNCH_COUNT:-560} --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 3 --profile-mode --no-graph > $out/b_ncu.log 2>&1;;
sanitize) timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > $out/memcheck.log 2>&1; echo "memcheck rc=$?" >> $out/memcheck.log; tail -4 $out/memcheck.log
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_ops_gpu.py -q -x -k "fps or ball or knn" > $out/racecheck.log 2>&1; echo "racecheck rc=$?" >> $out/racecheck.log; tail -4 $out/racecheck.log;;
smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/smoke.log 2>&1; tail -2 $out/smoke.log;;
ops) timeout 300 python tools/bench_ops.py > $out/bench_ops.log 2>&1; tail -3 $out/bench_ops.log; cp gpurun_out/bench_ops*.json $out/ 2>/dev/null;;
latency) timeout 300 python tools/bench_latency.py > $out/latency.log 2>&1; tail -3 $out/latency.log;;
gemm) export LD_LIBRARY_PATH=$PWD/hotrack_b200:$LD_LIBRARY_PATH; timeout 120 ./tools/dev/gemm_tc_check time > $out/gemm_check.log 2>&1; tail -3 $out/gemm_check.log;;
full) timeout 600 ncu --set full --clock-control none \
    -k regex:"${FULL_REGEX:-gemm_tc_kernel|wgrad_tc_kernel|cm_to_rows_bwd|rows_to_cm|pool_fwd|pool_bwd|sa_build_rows|fp_build_rows|sa_rows_bwd|fp_rows_bwd|fps_regs|ball_query|knn_kernel|three_nn}" \
    --launch-skip ${FULL_SKIP:-270} --launch-count ${FULL_COUNT:-90} -o $out/full -f \
    python bench.py --steps 1 --warmup 3 --profile-mode --no-graph > $out/full_ncu.log 2>&1
  ncu -i $out/full.ncu-rep --page raw --csv > $out/full_raw.csv 2>/dev/null
  sz=$(stat -c %s $out/full.ncu-rep); if [ "$sz" -gt 30000000 ]; then rm -f $out/full.ncu-rep; echo "rep dropped ($sz bytes)"; fi;;
esac; done
ls -la $out; du -sh gpurun_out
