#!/bin/bash
( time timeout 600 python -m pytest tests -q -m gpu --tb=short ) > gpurun_out/r1c_pytest_gpu.log 2>&1
tail -6 gpurun_out/r1c_pytest_gpu.log
for v in 0 1 2 0 1 2; do
  echo "SNERF_CG_EPILOGUE=$v"
  SNERF_CG_EPILOGUE=$v timeout 120 python tools/train_steps_tf32.py 8 2>&1 | tail -2
done > gpurun_out/r1c_train_ab.log 2>&1
cat gpurun_out/r1c_train_ab.log
