#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_stepfun.py tests/test_gpu_gridencoder.py -q -m gpu --tb=short > gpurun_out/pytest_stepfun.log 2>&1
tail -40 gpurun_out/pytest_stepfun.log
timeout 100 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_stepfun.py -q -m gpu -k "fixture or errors" --tb=line > gpurun_out/sanitizer_stepfun.log 2>&1
tail -5 gpurun_out/sanitizer_stepfun.log
timeout 100 python tools/stepfun_bench.py > gpurun_out/stepfun_bench.json 2>&1; cat gpurun_out/stepfun_bench.json
timeout 100 python tools/grid_bench.py --impl ours > gpurun_out/grid_bench_default.json 2>&1; tail -c 700 gpurun_out/grid_bench_default.json
