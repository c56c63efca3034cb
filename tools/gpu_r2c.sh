set -x
mkdir -p gpurun_out
nproc
timeout 600 python tests/big/time_dropin.py --reps 4 --verbose 3 --oracle-pairs 0 2>&1 | tee gpurun_out/r2_time_dropin_v3.log | tail -40
SONG_LOS_HOST_THREADS=16 timeout 600 python tests/big/time_dropin.py --reps 4 --verbose 3 --oracle-pairs 0 2>&1 | grep "pack\|transfer2_init #" | tail -9
SONG_LOS_HOST_THREADS=8 timeout 600 python tests/big/time_dropin.py --reps 4 --verbose 3 --oracle-pairs 0 2>&1 | grep "pack\|transfer2_init #" | tail -9
