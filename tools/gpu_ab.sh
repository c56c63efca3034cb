set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for v in 1 0; do
  SONG_LOS_FAST_VARIANT=$v timeout 600 python tools/time_configs.py --mode fast --reps 3 --oracle-pairs 30 sn_t_10percent_lmax2000 sn_pol_10percent_lmax2000 2>&1 | grep -E "FAST|oracle on"
done
