#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_stepfun.py -q -m gpu --tb=short > gpurun_out/pytest_stepfun.log 2>&1
tail -8 gpurun_out/pytest_stepfun.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:chan_gemm --launch-skip 58 -c 29 -f -o gpurun_out/prof_chan python tools/train_steps_tf32.py 3 > gpurun_out/ncu_chan.log 2>&1
tail -4 gpurun_out/ncu_chan.log
ls -la gpurun_out | tail -5
