# round 2: bench on configs[1] (dev), 1 GPU
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --workload sn_t_10percent_lmax2000 --steps 10 --warmup 3 --extra quick_song_run > gpurun_out/r2_bench_cfg1.json 2> gpurun_out/r2_bench_cfg1.err; tail -20 gpurun_out/r2_bench_cfg1.err; cat gpurun_out/r2_bench_cfg1.json
