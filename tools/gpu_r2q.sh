# e2e laps of transfer2_init on configs[3], one and two GPUs, plain process and under torchrun-like env
set -x
mkdir -p gpurun_out
SONG_LOS_DEVICES=1 timeout 900 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 3 --verbose 3 --oracle-pairs 0 > gpurun_out/r2q_dropin_cfg3_n1.log 2>&1
SONG_LOS_DEVICES=2 timeout 900 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 3 --verbose 3 --oracle-pairs 0 > gpurun_out/r2q_dropin_cfg3_n2.log 2>&1
SONG_LOS_DEVICES=2 OMP_NUM_THREADS=1 timeout 900 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 3 --verbose 2 --oracle-pairs 0 > gpurun_out/r2q_dropin_cfg3_n2_omp1.log 2>&1
grep "transfer2_init #" gpurun_out/r2q_dropin_cfg3_*.log
nproc; free -g | head -2
