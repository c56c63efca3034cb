# A/B of the DMMA kernel's consumer warp grid (SONG_LOS_TC_SHAPE=wm,wn,nw) on one workload
set -x
mkdir -p gpurun_out
for shape in ${SHAPES:-8,1,8 12,1,7 12,1,8 8,2,4 6,2,4}; do
  echo "=== shape $shape ${EXTRA_ENV}"
  env SONG_LOS_TC_SHAPE=$shape ${EXTRA_ENV} timeout 600 python tools/time_configs.py --mode fast --reps 3 --oracle-pairs 12 ${CONFIGS:-sn_t_10percent_lmax2000} 2>&1 | grep -E "FAST|Error|error" | tail -3
done
