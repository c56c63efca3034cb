set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_configs_gpu.py -m gpu -x -q -k row_index 2>&1 | tail -5
S=$(date +%s); timeout 2400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench wall: $(( $(date +%s) - S )) s"; tail -5 gpurun_out/r2_bench_default.err; cat gpurun_out/r2_bench_default.json
