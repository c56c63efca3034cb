# staged source upload in the drop-in (N=1), L2 eviction hints of the DMMA kernel's two operand streams (A/B: 128 = neutral, 64 = reversed)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
SONG_LOS_DEVICES=1 timeout 900 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 3 --verbose 3 --oracle-pairs 4 > gpurun_out/r2u_dropin_cfg3_n1.log 2>&1
grep "transfer2_init #\|max rel" gpurun_out/r2u_dropin_cfg3_n1.log
timeout 900 python tests/big/ab_knobs.py sn_pol_5percent_lmax3000 "SONG_LOS_TC_DBG=128;SONG_LOS_TC_DBG=64" --pair-stride 2 > gpurun_out/r2u_ab_l2_cfg3.log 2>&1; grep RESULT gpurun_out/r2u_ab_l2_cfg3.log
timeout 600 python tests/big/ab_knobs.py sn_t_10percent_lmax2000 "SONG_LOS_TC_DBG=128;SONG_LOS_TC_DBG=64" --reps 4 > gpurun_out/r2u_ab_l2_cfg1.log 2>&1; grep RESULT gpurun_out/r2u_ab_l2_cfg1.log
timeout 600 python tests/big/ab_knobs.py sn_pol_10percent_lmax2000 "SONG_LOS_TC_DBG=128" --reps 4 > gpurun_out/r2u_ab_l2_cfg2.log 2>&1; grep RESULT gpurun_out/r2u_ab_l2_cfg2.log
