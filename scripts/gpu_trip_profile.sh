#!/bin/bash
# ncu evidence for the current kernels: launch list of one layer + full captures of the top kernels.
set +e
mkdir -p gpurun_out
CMD="python bench.py --gpus 1 --steps 1 --warmup 0 --layers 1 --no-e2e --no-cpu-baseline --no-shared --streams 1"
timeout 300 $CMD > gpurun_out/plain_prof.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_prof.log; exit 1; }
echo "=== launch list"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r01b.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "exit $?"
echo "=== full: hessian_tc (7 launches = one per linear)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hessian_tc -c 7 -o gpurun_out/prof_hessian_tc_r01b -f $CMD > gpurun_out/ncu_h.log 2>&1; echo "exit $?"
echo "=== full: gemm_tf32x3 (feedback + cholesky)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 100 -c 4 -o gpurun_out/prof_gemm_tc_r01b -f $CMD > gpurun_out/ncu_g.log 2>&1; echo "exit $?"
echo "=== full: atq_block, ssr, pack-free kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"atq_block|ssr_colstats|ssr_rowmean|ssr_select|feedback_coef|chol_diag" -s 60 -c 12 -o gpurun_out/prof_sweep_misc_r01b -f $CMD > gpurun_out/ncu_m.log 2>&1; echo "exit $?"
ls -la gpurun_out | grep -E "r01b|launches"
