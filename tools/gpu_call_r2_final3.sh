mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/r2g_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2g_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2g_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_r2.py > gpurun_out/r2g_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/r2g_sanitizer.log
SNERF_B200_PAIR=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_r2.py > gpurun_out/r2g_sanitizer_pair.log 2>&1; echo "sanitizer(pair) rc=$?"; tail -4 gpurun_out/r2g_sanitizer_pair.log
