set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "=== $*"; env "$@" timeout 600 python tools/time_configs.py --mode fast --reps 3 --oracle-pairs 6 ${CONFIGS:-sn_t_10percent_lmax2000} 2>&1 | grep -E "FAST|Error|error" | tail -2 | cut -c1-200; }
CONFIGS=sn_pol_10percent_lmax2000 run A=1
run A=1
