#!/bin/bash
# Closing validation of the round (one B200):  gpurun --timeout 2400 -- 'bash tools/gpu_call_r2_final5.sh'
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r2j_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2j_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2j_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_r2.py > gpurun_out/r2j_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/r2j_sanitizer.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err; tail -c 300 gpurun_out/r2j_bench_ref.json
timeout 900 python bench.py > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err; tail -c 1500 gpurun_out/r2j_bench_n1.json; tail -2 gpurun_out/r2j_bench_n1.err
