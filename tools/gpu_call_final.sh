#!/bin/bash
# round-end style verification: all GPU tests, smoke, both bench arms
( time timeout 600 python -m pytest tests -q -m gpu --durations=5 --tb=short ) > gpurun_out/r1c_pytest_gpu.log 2>&1
tail -12 gpurun_out/r1c_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1c_smoke.log 2>&1; tail -6 gpurun_out/r1c_smoke.log
( time timeout 500 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r1c_bench_ref.json 2> gpurun_out/r1c_bench_ref.err; tail -c 600 gpurun_out/r1c_bench_ref.json; tail -3 gpurun_out/r1c_bench_ref.err
( time timeout 800 python bench.py ) > gpurun_out/r1c_bench_n1.json 2> gpurun_out/r1c_bench_n1.err; tail -c 1500 gpurun_out/r1c_bench_n1.json; tail -4 gpurun_out/r1c_bench_n1.err
