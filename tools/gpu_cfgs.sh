set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in ${CONFIGS:-quick_song_run sn_t_10percent_lmax2000 sn_pol_10percent_lmax2000}; do
  timeout 1500 python tools/run_config.py $c --oracle-pairs ${ORACLE_PAIRS:-8} > gpurun_out/cfg_$c.json 2> gpurun_out/cfg_$c.err; tail -4 gpurun_out/cfg_$c.err; cat gpurun_out/cfg_$c.json
done
