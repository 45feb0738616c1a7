#!/bin/bash
# round-2 ncu artefacts: launch list of one transformer layer through bench.py + --set full of the feedback GEMM and the
# diagonal-block kernel inside a 4096 x 11008 chain.  Every profiled command first runs plain and must exit 0.
set +e
mkdir -p gpurun_out
CMD="python bench.py --gpus 1 --steps 1 --warmup 1 --layers 1 --no-e2e --no-cpu-baseline --no-shared --no-packed --no-ref-cuda --no-whole-model"
timeout 300 $CMD > gpurun_out/plain_l1.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_l1.log; exit 1; }
echo "=== launch list (second pass over the layer = the timed step)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "exit $?"
export N=4096 M=11008 NT=16384 REPS=1
python scripts/one_linear.py > gpurun_out/plain_one.log 2>&1 || { echo "plain one_linear failed"; exit 1; }
echo "=== full: feedback GEMM (gathered + statistics), two launches mid-sweep"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 300 -c 2 -o gpurun_out/prof_r02_gemm_fb_final -f python scripts/one_linear.py > gpurun_out/ncu_b.log 2>&1; echo "exit $?"
echo "=== full: chol_diag"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_diag -s 40 -c 2 -o gpurun_out/prof_r02_diag_final -f python scripts/one_linear.py > gpurun_out/ncu_c.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02.csv
