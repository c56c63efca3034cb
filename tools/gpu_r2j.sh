set -x
mkdir -p gpurun_out
S1="SONG_LOS_TC_SHAPE=8/1/8;SONG_LOS_TC_SHAPE=4/2/8;SONG_LOS_TC_SHAPE=2/4/8;SONG_LOS_TC_SHAPE=6/2/5;SONG_LOS_TC_RMAX=6;SONG_LOS_TC_RMAX=8;SONG_LOS_TC_SHAPE=8/1/8,SONG_LOS_TC_RMAX=6;SONG_LOS_TC_SHAPE=4/2/8,SONG_LOS_TC_RMAX=4;SONG_LOS_TC_SHAPE=4/2/8,SONG_LOS_TC_RMAX=6"
for stride in 2 4 8 16; do
  timeout 600 python tests/big/ab_knobs.py sn_t_10percent_lmax2000 "$S1" --pair-stride $stride --reps 3 2>&1 | grep -E "RESULT|pairs,|Error" | tee -a gpurun_out/r2_ab_density_cfg1.log
done
S3="SONG_LOS_TC_SHAPE=12/1/7;SONG_LOS_TC_SHAPE=4/2/8;SONG_LOS_TC_SHAPE=2/4/8;SONG_LOS_TC_RMAX=4;SONG_LOS_TC_SHAPE=12/1/7,SONG_LOS_TC_RMAX=6;SONG_LOS_TC_SHAPE=6/2/8;SONG_LOS_TC_SHAPE=8/1/8"
for stride in 4 8; do
  timeout 900 python tests/big/ab_knobs.py sn_pol_5percent_lmax3000 "$S3" --pair-stride $stride --reps 2 2>&1 | grep -E "RESULT|pairs,|Error" | tee -a gpurun_out/r2_ab_density_cfg3.log
done
