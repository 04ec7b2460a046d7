#!/bin/bash
# First GPU call of the next round (one B200, ~6 min of box time): everything the round-1 budget no longer covered.
#   gpurun --timeout 900 -- 'bash tools/gpu_call_round2_first.sh'
mkdir -p gpurun_out
# 1. parity + smoke on the current tree
( time timeout 400 python -m pytest tests -q -m gpu --tb=short --durations=5 ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -6 gpurun_out/r2_smoke.log
# 2. both bench arms (the chunk32768 row was added after the last GPU run of round 1: first numbers here)
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 900 gpurun_out/r2_bench_n1.json
# 3. launch list of the bench command (durations only; shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1; wc -l gpurun_out/r2_launches_bench.csv
# 4. full captures of the training kernels that have none yet (one launch each, third step)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"heads_fwd_kernel|head_bwd_kernel|dw_tf32_kernel|skinny_kernel|loss_fwd_kernel|pack_jobs_kernel" \
    --launch-skip 16 -c 8 -f -o gpurun_out/r2_prof_train_small python tools/train_steps_tf32.py 3 > gpurun_out/r2_ncu_train_small.log 2>&1; tail -2 gpurun_out/r2_ncu_train_small.log
ls -la gpurun_out | tail -12
