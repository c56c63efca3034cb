# drop-in at N=1: error flags no longer through the copy engine, four ramped batches
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SONG_LOS_DEVICES=1 timeout 900 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 4 --verbose 3 --oracle-pairs 4 > gpurun_out/r2x_dropin_cfg3_n1.log 2>&1
grep "transfer2_init #\|max rel\|pairs in" gpurun_out/r2x_dropin_cfg3_n1.log
grep "batch of" gpurun_out/r2x_dropin_cfg3_n1.log | tail -8
SONG_LOS_MIN_BATCHES=3 SONG_LOS_DEVICES=1 timeout 900 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 4 --verbose 3 --oracle-pairs 0 > gpurun_out/r2x_dropin_cfg3_n1_b3.log 2>&1
grep "transfer2_init #\|max rel\|pairs in" gpurun_out/r2x_dropin_cfg3_n1_b3.log
grep "batch of" gpurun_out/r2x_dropin_cfg3_n1_b3.log | tail -6
