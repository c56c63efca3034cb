# round 2 final artefacts at N=1: GPU suite, default bench, reference arm, ncu launch list + full capture of the dominant kernel
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
S=$(date +%s); timeout 2400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err; echo "bench wall: $(( $(date +%s) - S )) s"; tail -3 gpurun_out/r2_final_bench_1gpu.err; cut -c1-1200 gpurun_out/r2_final_bench_1gpu.json
S=$(date +%s); timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err; echo "reference wall: $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/r2_final_bench_ref.json
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_final_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --extra none --e2e-steps 0 --no-cpu-baseline --no-k3-integral > gpurun_out/r2_final_bench_under_ncu.log 2>&1; tail -1 gpurun_out/r2_final_bench_under_ncu.log | cut -c1-200
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:los_tc_kernel -s 12 -c 1 -f -o gpurun_out/r2_final_prof_tc_cfg3 python bench.py --steps 1 --warmup 3 --extra none --e2e-steps 0 --no-cpu-baseline --no-k3-integral > gpurun_out/r2_final_ncu_full.log 2>&1; tail -2 gpurun_out/r2_final_ncu_full.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
