#!/bin/bash
# Round-2 closing evidence (one B200):  gpurun --timeout 2400 -- 'bash tools/gpu_call_r2_final.sh'
mkdir -p gpurun_out
( time timeout 700 python -m pytest tests -q -m gpu --tb=short --durations=5 ) > gpurun_out/r2f_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -8 gpurun_out/r2f_smoke.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -c 1800 gpurun_out/r2f_bench_n1.json
# launch list of the WHOLE bench command (durations only; shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 5 > gpurun_out/r2f_ncu_bench.log 2>&1; wc -l gpurun_out/r2f_launches_bench.csv
ls -la gpurun_out | tail -8
# memcheck over small invocations of every round-2 kernel
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_r2.py > gpurun_out/r2f_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/r2f_sanitizer.log
