set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "=== $*"; env "$@" timeout 600 python tools/time_configs.py --mode fast --reps 3 --oracle-pairs 6 ${CONFIGS:-sn_t_10percent_lmax2000} 2>&1 | grep -E "FAST|Error|error" | tail -2 | cut -c1-200; }
run A=1
run SONG_LOS_TC_NP1=1
run SONG_LOS_TC_SHAPE=12,1,8
run SONG_LOS_TC_SHAPE=8,1,8
CONFIGS=sn_pol_10percent_lmax2000 run A=1
