#!/bin/bash
( time timeout 600 python -m pytest tests -q -m gpu --tb=short ) > gpurun_out/r1c_pytest_gpu.log 2>&1
tail -8 gpurun_out/r1c_pytest_gpu.log
timeout 120 python tools/train_steps_tf32.py 8 2>&1 | tail -3 > gpurun_out/r1c_train_ab.log 2>&1
cat gpurun_out/r1c_train_ab.log
( time timeout 800 python bench.py ) > gpurun_out/r1c_bench_n1.json 2> gpurun_out/r1c_bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1c_bench_n1.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'clocks',d['clocks'])
print('train',{k:d['train'][k] for k in ('value','ms_per_step','tflops_per_gpu','final_loss')}, d['train']['e2e'], d['train']['cpu_baseline'])
PY
tail -3 gpurun_out/r1c_bench_n1.err
