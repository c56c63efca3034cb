set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropin.py -m gpu -x -q 2>&1 | tail -3
SONG_LOS_DEVICES=1 timeout 900 python tests/big/time_dropin.py --reps 5 --verbose 0 --oracle-pairs 0 2>&1 | grep "transfer2_init #"
SONG_LOS_DEVICES=2 timeout 900 python tests/big/time_dropin.py --reps 5 --verbose 3 --oracle-pairs 0 2>&1 | grep -E "transfer2_init #|batches|done at|submitted"
SONG_LOS_DEVICES=2 timeout 1500 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 4 --verbose 3 --oracle-pairs 4 2>&1 | grep -E "transfer2_init #|batches|done at|submitted|max rel|harness"
SONG_LOS_DEVICES=1 timeout 1500 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 4 --verbose 3 --oracle-pairs 0 2>&1 | grep -E "transfer2_init #|batches|done at|submitted|max rel|harness"
