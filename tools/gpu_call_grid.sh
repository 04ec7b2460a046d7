#!/bin/bash
# one GPU-box call: golden fixtures from the reference's grid kernels, grid parity tests, micro-benchmark, ncu captures
mkdir -p gpurun_out/golden
python oracle/make_golden_grid.py gpurun_out/golden > gpurun_out/golden.log 2>&1 && cp gpurun_out/golden/*.npz tests/golden/
timeout 300 python -m pytest tests/test_gpu_gridencoder.py -q -m gpu --tb=short > gpurun_out/pytest_grid.log 2>&1
tail -30 gpurun_out/pytest_grid.log
timeout 200 python tools/grid_bench.py --impl both > gpurun_out/grid_bench.jsonl 2>&1
cat gpurun_out/grid_bench.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:grid_ --launch-skip 3 -c 3 -f -o gpurun_out/prof_grid python tools/grid_bench.py --impl ours --steps 1 > gpurun_out/ncu_grid.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:kernel_grid --launch-skip 3 -c 3 -f -o gpurun_out/prof_grid_ref python tools/grid_bench.py --impl reference --steps 1 > gpurun_out/ncu_grid_ref.log 2>&1
ls -la gpurun_out
