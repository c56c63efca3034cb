# one ncu --set full capture of the LOS kernel on the bench workload (+ stall-sampled SASS)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:los_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
