#!/bin/bash
set +e
mkdir -p gpurun_out
echo "=== bench N=1"
timeout 1200 python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "exit $?"
tail -n 3 gpurun_out/bench_n1.log; tail -n 5 gpurun_out/bench_n1.err
echo "=== ncu launch list (1 layer)"
CMD="python bench.py --gpus 1 --steps 1 --warmup 1 --layers 1 --no-e2e --no-cpu-baseline --no-shared"
timeout 600 $CMD > gpurun_out/plain_1layer.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/plain_1layer.log | cut -c1-400
echo "=== ncu full: hessian_tc"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hessian_tc -s 7 -c 2 -o gpurun_out/prof_hessian_tc_r01 -f $CMD > gpurun_out/ncu_full_h.log 2>&1; echo "exit $?"
ls -la gpurun_out | head -30
