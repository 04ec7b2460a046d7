#!/bin/bash
# one GPU-box call: golden fixtures from the reference's grid kernels, grid parity tests, micro-benchmark, ncu captures
mkdir -p gpurun_out/golden
python oracle/make_golden_grid.py gpurun_out/golden > gpurun_out/golden.log 2>&1; cat gpurun_out/golden.log | tail -5
cp gpurun_out/golden/*.npz tests/golden/ 2>/dev/null
timeout 400 python -m pytest tests/test_gpu_gridencoder.py -q -m gpu --tb=short > gpurun_out/pytest_grid.log 2>&1
tail -40 gpurun_out/pytest_grid.log
: > gpurun_out/grid_bench.jsonl
for v in 0 1 2 3 4 8; do
  echo "variant $v" >> gpurun_out/grid_bench.jsonl
  SNERF_GRID_VARIANT=$v timeout 200 python tools/grid_bench.py --impl ours >> gpurun_out/grid_bench.jsonl 2>&1
done
timeout 200 python tools/grid_bench.py --impl reference >> gpurun_out/grid_bench.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/grid_bench.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k.endswith('_ms') or k=='impl'})
    else: print(l.strip())
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"grid_(ms_)?(fwd|bwd)2?_kernel" --launch-skip 4 -c 8 -f -o gpurun_out/prof_grid python tools/grid_bench.py --impl ours --steps 1 --warmup 1 > gpurun_out/ncu_grid.log 2>&1
tail -3 gpurun_out/ncu_grid.log
ls -la gpurun_out | head -40
