# every BASELINE workload end to end through the pipelined C-ABI, with sampled oracle parity
set -x
mkdir -p gpurun_out
for c in ${CONFIGS:-quick_song_run sn_t_10percent_lmax2000 sn_pol_10percent_lmax2000 sn_pol_5percent_lmax3000}; do
  timeout 1500 python tests/big/run_config.py $c --oracle-pairs ${ORACLE_PAIRS:-8} > gpurun_out/cfg_$c.json 2> gpurun_out/cfg_$c.err; tail -2 gpurun_out/cfg_$c.err; cat gpurun_out/cfg_$c.json
done
# configs[4] is quoted at 8 GPUs: one rank's shard (every 8th (k1,k2) pair) on this GPU
timeout 1500 python tests/big/run_config.py sn_pol_5percent_lmax4000 --pair-stride 8 --oracle-pairs 4 > gpurun_out/cfg_sn_pol_5percent_lmax4000_shard8.json 2> gpurun_out/cfg_sn_pol_5percent_lmax4000_shard8.err; tail -2 gpurun_out/cfg_sn_pol_5percent_lmax4000_shard8.err; cat gpurun_out/cfg_sn_pol_5percent_lmax4000_shard8.json
