set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python tools/time_configs.py --mode fast --reps 3 --oracle-pairs 30 ${CONFIGS:-sn_t_10percent_lmax2000 sn_pol_10percent_lmax2000} 2>&1 | grep -E "FAST|oracle on|Error|error"
