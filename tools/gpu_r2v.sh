# drop-in at N=1: three ramped batches; exposed D2H tail with and without the early block-wise copy
set -x
mkdir -p gpurun_out
SONG_LOS_DEVICES=1 timeout 900 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 3 --verbose 3 --oracle-pairs 4 > gpurun_out/r2v_dropin_cfg3_n1.log 2>&1
SONG_LOS_NO_EARLY_D2H=1 SONG_LOS_DEVICES=1 timeout 900 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 3 --verbose 3 --oracle-pairs 0 > gpurun_out/r2v_dropin_cfg3_n1_noearly.log 2>&1
grep "transfer2_init #\|max rel" gpurun_out/r2v_dropin_cfg3_n1*.log
grep "batch of\|since" gpurun_out/r2v_dropin_cfg3_n1.log | tail -16
grep "batch of" gpurun_out/r2v_dropin_cfg3_n1_noearly.log | tail -6
