set -x
mkdir -p gpurun_out
nproc
timeout 900 python bench.py --workload sn_t_10percent_lmax2000 --steps 5 --warmup 3 --extra sn_t_10percent_lmax2000 --no-cpu-baseline --dropin-verbose 3 2>&1 | grep -v '^{"metric"' | tail -60
