set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { echo "=== $*"; env "$@" timeout 600 python tools/time_configs.py --mode fast --reps 2 --oracle-pairs 6 ${CONFIGS:-sn_t_10percent_lmax2000} 2>&1 | grep -E "FAST|Error|error" | tail -2 | cut -c1-200; }
run A=1
CONFIGS=sn_pol_10percent_lmax2000 run A=1
run3() { echo "=== $*"; env "$@" timeout 900 python tools/run_config.py sn_pol_5percent_lmax3000 "${EXTRA[@]}" --oracle-pairs 2 > gpurun_out/r3.out 2> gpurun_out/r3.err; tail -1 gpurun_out/r3.err | cut -c1-300; python tools/print_r3.py; }
EXTRA=(--pair-stride 5); run3 A=1
EXTRA=(--pair-stride 2); run3 A=1
