#!/bin/bash
# First GPU trip: every parity group in its own process (a faulting kernel only takes its own group down).
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest -q -p no:cacheprovider -m gpu "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 15 gpurun_out/$name.log; }
run atq      tests/test_gpu_kernels.py -k "atq"
run ssr      tests/test_gpu_kernels.py -k "ssr"
run hffma    tests/test_gpu_kernels.py -k "hessian_ffma"
run htc      tests/test_gpu_kernels.py -k "hessian_tcgen05 or hessian_auto"
run chol     tests/test_gpu_kernels.py -k "inverse or cholesky"
run fb       tests/test_gpu_kernels.py -k "feedback"
run codec    tests/test_gpu_kernels.py -k "pack"
run layer    tests/test_gpu_layer.py -s
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -n 8 gpurun_out/smoke.log
