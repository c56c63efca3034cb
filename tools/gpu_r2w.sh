# drop-in at N=1 after the capacity fix: three ramped batches, steady state after the first call
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SONG_LOS_DEVICES=1 timeout 900 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 4 --verbose 3 --oracle-pairs 4 > gpurun_out/r2w_dropin_cfg3_n1.log 2>&1
grep "transfer2_init #\|max rel\|pairs in" gpurun_out/r2w_dropin_cfg3_n1.log
grep "batch of\|since" gpurun_out/r2w_dropin_cfg3_n1.log | tail -14
timeout 600 python tests/big/time_dropin.py sn_t_10percent_lmax2000 --reps 5 --verbose 3 --oracle-pairs 4 > gpurun_out/r2w_dropin_cfg1_n1.log 2>&1
grep "transfer2_init #\|max rel\|pairs in" gpurun_out/r2w_dropin_cfg1_n1.log
