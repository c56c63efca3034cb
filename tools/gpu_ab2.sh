set -x
mkdir -p gpurun_out
run() { echo "=== $*"; env "$@" timeout 600 python tools/time_configs.py --mode fast --reps 3 --oracle-pairs 6 ${CONFIGS:-sn_t_10percent_lmax2000} 2>&1 | grep -E "FAST|Error|error" | tail -2 | head -1 | cut -c1-120; }
run A=1
run SONG_LOS_TC_TAUC=32
run SONG_LOS_TC_TAUC=20
run SONG_LOS_TC_TAUC=16
run SONG_LOS_TC_TAUC=12
