#!/bin/bash
# GPU trip for the packed ternary layer: parity tests, bandwidth table, ncu captures.  Every step has
# its own timeout and log under gpurun_out/; the fused tcgen05 GEMM runs in its own processes (a trap there must not
# take the other results down).
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== tl tests";   timeout 300 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_ternary_linear.py > gpurun_out/tl_tests.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/tl_tests.log
echo "=== fused tests"; timeout 240 python -m pytest -q -p no:cacheprovider -m gpu tests/test_gpu_ternary_gemm.py > gpurun_out/tl_fused_tests.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/tl_fused_tests.log
echo "=== tl bench";   TL_BENCH_PREFILL=0 timeout 300 python scripts/tl_bench.py > gpurun_out/tl_bench.log 2>&1; echo "exit $?"; tail -n 30 gpurun_out/tl_bench.log
echo "=== prefill bench"; TL_BENCH_PREFILL=only timeout 200 python scripts/tl_bench.py > gpurun_out/tl_bench_prefill.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/tl_bench_prefill.log
echo "=== ncu full: tl kernels"
timeout 120 python scripts/tl_profile_case.py > gpurun_out/tl_case.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tl_gemv -c 4 -o gpurun_out/prof_tl_gemv_r01d -f python scripts/tl_profile_case.py > gpurun_out/ncu_tl_gemv.log 2>&1
echo "exit $?"
TL_FUSED=1 timeout 120 python scripts/tl_profile_case.py > gpurun_out/tl_case_fused.log 2>&1 && \
TL_FUSED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tl_gemm_tc|tl_expand" -c 4 -o gpurun_out/prof_tl_gemm_r01d -f python scripts/tl_profile_case.py > gpurun_out/ncu_tl_gemm.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/tl_case_fused.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/smoke.log
ls -la gpurun_out | head -40
