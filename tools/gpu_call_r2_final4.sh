#!/bin/bash
# Round-2 closing evidence call (one B200):  gpurun --timeout 2400 -- 'bash tools/gpu_call_r2_final4.sh'
mkdir -p gpurun_out
FLAGS="--no-cpu-baseline --no-train --no-parity-mode --no-frame6 --no-grid --no-mip"
# 1. both bench arms, default invocation (what the driver runs)
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_bench_ref.json 2> gpurun_out/r2i_bench_ref.err; tail -c 400 gpurun_out/r2i_bench_ref.json
timeout 900 python bench.py > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; tail -c 1200 gpurun_out/r2i_bench_n1.json; tail -2 gpurun_out/r2i_bench_n1.err
# 2. launch list of the bench command (durations only: shares of the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2i_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2i_ncu_launches.log 2>&1; wc -l gpurun_out/r2i_launches_bench.csv
# 3. full capture of the headline kernel at the headline problem size (DRAM traffic per launch, tensor pipe, L2->SM bytes)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:snerf_bf16_render_kernel -s 3 -c 1 -f -o gpurun_out/r2i_fused_full \
    python bench.py --steps 1 --warmup 1 $FLAGS > gpurun_out/r2i_ncu_full.log 2>&1; tail -2 gpurun_out/r2i_ncu_full.log
# 4. training-step kernels after the forward's epilogue change (third step)
timeout 600 ncu --set full --clock-control none -k regex:"dx_chain_tc_kernel|dw_tc_kernel|snerf_bf16_render_kernel" \
    --launch-skip 6 -c 3 -f -o gpurun_out/r2i_prof_train_tc python tools/train_steps.py 3 512 bf16 > gpurun_out/r2i_ncu_train_tc.log 2>&1; tail -2 gpurun_out/r2i_ncu_train_tc.log
ls -la gpurun_out | tail -8
