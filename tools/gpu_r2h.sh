# round 2: ncu launch list of the bench step + one full capture of the dominant kernel on the default workload (configs[3])
set -x
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --extra none --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1; tail -2 gpurun_out/r2_bench_under_ncu.log | cut -c1-300
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:los_tc_kernel -s 6 -c 2 -f -o gpurun_out/r2_prof_tc_cfg3 python bench.py --steps 1 --warmup 3 --extra none --e2e-steps 0 --no-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1; tail -3 gpurun_out/r2_ncu_full.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
