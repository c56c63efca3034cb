# ncu --set full capture of the DMMA kernel launches of one workload (default sn_pol_10percent_lmax2000: T block, E block)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:los_tc_kernel -s ${SKIP:-2} -c ${COUNT:-2} -f -o gpurun_out/prof_cfg python tests/big/time_configs.py --mode fast --reps 2 --oracle-pairs 2 ${CONFIGS:-sn_pol_10percent_lmax2000} > gpurun_out/ncu_cfg.log 2>&1; tail -3 gpurun_out/ncu_cfg.log
