# round 2, first GPU visit: GPU tests + wall time of the drop-in
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 600 python tests/big/time_dropin.py --reps 5 2>&1 | tee gpurun_out/r2_time_dropin.log | tail -60
SONG_LOS_MIN_BATCHES=1 timeout 600 python tests/big/time_dropin.py --reps 4 --verbose 0 --oracle-pairs 0 2>&1 | grep transfer2_init | tee gpurun_out/r2_time_dropin_b1.log
SONG_LOS_MIN_BATCHES=2 timeout 600 python tests/big/time_dropin.py --reps 4 --verbose 0 --oracle-pairs 0 2>&1 | grep transfer2_init | tee gpurun_out/r2_time_dropin_b2.log
SONG_LOS_MIN_BATCHES=8 timeout 600 python tests/big/time_dropin.py --reps 4 --verbose 0 --oracle-pairs 0 2>&1 | grep transfer2_init | tee gpurun_out/r2_time_dropin_b8.log
timeout 600 python tests/big/time_dropin.py --reps 3 --ref-bessel2 --oracle-pairs 0 2>&1 | tee gpurun_out/r2_time_dropin_refb2.log | tail -30
