# compute-sanitizer memcheck over a subset of the GPU tests (tiny workloads)
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file gpurun_out/memcheck.log python -m pytest tests/test_parity_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -k "${K:-tiny or chunks or pipeline_is_bitwise or time_grids or temperature}" 2>&1 | tail -5
echo "exit: $?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/memcheck.log | head -20
