# A/B harness for kernel experiments on the B200 (used for every measurement quoted in DESIGN.md section 4):
#   run  ENV=VAL...    -> times song_los_compute(FAST) on ${CONFIGS} (default configs[1]) with a sampled oracle check
#   run3 ENV=VAL...    -> sn_pol_5percent_lmax3000, every 5th pair (1.5 M k3 values, one batch), through tests/big/run_config.py
# knobs: SONG_LOS_TC_SHAPE=wm,wn,nw  SONG_LOS_TC_KCS  SONG_LOS_TC_RMAX  SONG_LOS_TC_STAGES  SONG_LOS_TC_TAUC
#        SONG_LOS_TC_DBG (4 no MMAs, 8 no source copies, 16 no table copies, 32 no header prefetch; results invalid)
set -x
mkdir -p gpurun_out
run() { echo "=== $*"; env "$@" timeout 600 python tests/big/time_configs.py --mode fast --reps 3 --oracle-pairs 6 ${CONFIGS:-sn_t_10percent_lmax2000} 2>&1 | grep -E "FAST|Error|error" | tail -2 | cut -c1-200; }
run3() { echo "=== $*"; env "$@" timeout 900 python tests/big/run_config.py sn_pol_5percent_lmax3000 --pair-stride 5 --oracle-pairs 2 2> gpurun_out/r3.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('RESULT', d['k3_units'], d['batches'], d['gpu_ms'], d['parity']['max_rel_err_above_1e-6_of_max'])"; }
run A=1
CONFIGS=sn_pol_10percent_lmax2000 run A=1
run3 A=1
