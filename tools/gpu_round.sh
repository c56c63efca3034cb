# one GPU visit: sanitizer on the smoke test, the GPU test suite, timings
set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1; echo "sanitizer exit $?"; tail -5 gpurun_out/sanitizer.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python tools/time_configs.py --mode both quick_song_run sn_t_10percent_lmax2000 2>&1 | tail -40
timeout 900 python tools/time_configs.py --mode fast sn_pol_10percent_lmax2000 2>&1 | tail -12
