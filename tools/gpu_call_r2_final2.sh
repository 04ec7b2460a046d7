mkdir -p gpurun_out
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; tail -c 600 gpurun_out/r2f_bench_ref.json; tail -3 gpurun_out/r2f_bench_ref.err
timeout 900 python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -c 1500 gpurun_out/r2f_bench_n1.json; tail -3 gpurun_out/r2f_bench_n1.err
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_train_tc.py -q -s -k "convergence or graphed" 2>&1 | grep -E "^\[|passed|failed"; done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_r2.py > gpurun_out/r2f_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/r2f_sanitizer.log
