# round 2: full GPU suite, sanitizers on the FAST path, default bench (configs[3]) at N=1
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 77 --log-file gpurun_out/r2_racecheck.log python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fast_tolerance_tiny or several_k_chunks" 2>&1 | tail -3
echo "racecheck exit: $?"; grep -E "RACECHECK SUMMARY|hazard|ERROR SUMMARY" gpurun_out/r2_racecheck.log | sort | uniq -c | head -10
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 77 --log-file gpurun_out/r2_synccheck.log python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "fast_tolerance_tiny or several_k_chunks" 2>&1 | tail -3
echo "synccheck exit: $?"; grep -E "ERROR SUMMARY|Barrier|divergent" gpurun_out/r2_synccheck.log | sort | uniq -c | head -10
/usr/bin/time -v timeout 2400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; grep -E "Elapsed|Maximum resident" gpurun_out/r2_bench_default.err; tail -5 gpurun_out/r2_bench_default.err; cat gpurun_out/r2_bench_default.json
