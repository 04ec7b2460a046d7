#!/bin/bash
# Round-2 evidence call (one B200):  gpurun --timeout 2400 -- 'bash tools/gpu_call_r2_capture.sh'
mkdir -p gpurun_out
# 1. parity + smoke on the current tree
( time timeout 600 python -m pytest tests -q -m gpu --tb=short --durations=5 ) > gpurun_out/r2_pytest_gpu.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -6 gpurun_out/r2_smoke.log
# 2. both bench arms
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 1500 gpurun_out/r2_bench_n1.json
# 3. launch lists (durations only; shares, not absolutes): the training step and the mip forward
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_train_bf16_launches.csv \
    python tools/train_steps.py 3 512 bf16 > gpurun_out/r2_ncu_train.log 2>&1; wc -l gpurun_out/r2_train_bf16_launches.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_mip_launches.csv \
    python tools/mip_bench.py 2048 > gpurun_out/r2_ncu_mip.log 2>&1; wc -l gpurun_out/r2_mip_launches.csv
# 4. full captures: the kernels of the tensor-core training step (third step) and one 1024-wide layer GEMM of the mip path
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dx_chain_tc_kernel|dw_tc_kernel|snerf_bf16_render_kernel|composite_bwd_kernel|adam_kernel" \
    --launch-skip 10 -c 5 -f -o gpurun_out/r2_prof_train_tc python tools/train_steps.py 3 512 bf16 > gpurun_out/r2_ncu_train_tc.log 2>&1; tail -2 gpurun_out/r2_ncu_train_tc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lin_tc_kernel" --launch-skip 53 -c 2 -f -o gpurun_out/r2_prof_lin_tc \
    python tools/mip_bench.py 2048 > gpurun_out/r2_ncu_lin.log 2>&1; tail -2 gpurun_out/r2_ncu_lin.log
ls -la gpurun_out | tail -14
