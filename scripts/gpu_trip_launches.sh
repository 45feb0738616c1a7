#!/bin/bash
# ncu launch list (durations only) of one transformer layer through bench.py
set +e
mkdir -p gpurun_out
TAG=${1:-r01c}
CMD="python bench.py --gpus 1 --steps 1 --warmup 0 --layers 1 --no-e2e --no-cpu-baseline --no-shared --streams 1"
timeout 300 $CMD > gpurun_out/plain_prof.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_prof.log; exit 1; }
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l.log 2>&1; echo "exit $?"
python scripts/summarize_ncu.py launches gpurun_out/launches_$TAG.csv | tee gpurun_out/launches_$TAG.txt
