set -x
mkdir -p gpurun_out
run3() { echo "=== $*"; env "$@" timeout 900 python tools/run_config.py sn_pol_5percent_lmax3000 --pair-stride 5 --oracle-pairs 2 > gpurun_out/r3.out 2> gpurun_out/r3.err; tail -1 gpurun_out/r3.err | cut -c1-300; python tools/print_r3.py; }
cat > tools/print_r3.py <<'PY'
import json
d=json.load(open('gpurun_out/r3.out')); print('RESULT', d['gpu_ms']['ms_integrate'], d['gpu_ms']['ms_total'], d['parity']['max_rel_err_above_1e-6_of_max'])
PY
run3 A=1
run3 SONG_LOS_TC_KCS=3
run3 SONG_LOS_TC_KCS=4
run3 SONG_LOS_TC_KCS=2
run3 SONG_LOS_TC_SHAPE=6,2,8
run3 SONG_LOS_TC_SHAPE=6,2,8 SONG_LOS_TC_KCS=3
run3 SONG_LOS_TC_SHAPE=12,1,8 SONG_LOS_TC_RMAX=7
