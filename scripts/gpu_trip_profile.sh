#!/bin/bash
# ncu --set full captures of the top kernels (one transformer layer through bench.py); launch list: gpu_trip_launches.sh
set +e
mkdir -p gpurun_out
TAG=${1:-r01c}
CMD="python bench.py --gpus 1 --steps 1 --warmup 0 --layers 1 --no-e2e --no-cpu-baseline --no-shared --streams 1"
timeout 300 $CMD > gpurun_out/plain_prof.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_prof.log; exit 1; }
echo "=== full: hessian_tc (7 launches = one per linear)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hessian_tc -c 7 -o gpurun_out/prof_hessian_tc_$TAG -f $CMD > gpurun_out/ncu_h.log 2>&1; echo "exit $?"
echo "=== full: gemm_tf32x3 (feedback + cholesky)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 100 -c 4 -o gpurun_out/prof_gemm_tc_$TAG -f $CMD > gpurun_out/ncu_g.log 2>&1; echo "exit $?"
echo "=== full: atq_block, ssr, coef, diag kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"atq_block|ssr_sims|ssr_select|feedback_coef|chol_diag|aga_vector" -s 60 -c 12 -o gpurun_out/prof_sweep_misc_$TAG -f $CMD > gpurun_out/ncu_m.log 2>&1; echo "exit $?"
ls -la gpurun_out | grep -E "$TAG"
