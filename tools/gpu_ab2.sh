set -x
mkdir -p gpurun_out
run3() { echo "=== $*"; env "$@" timeout 900 python tools/run_config.py sn_pol_5percent_lmax3000 --pair-stride 5 --oracle-pairs 2 > gpurun_out/r3.out 2> gpurun_out/r3.err; tail -2 gpurun_out/r3.err | cut -c1-300; python -c "
import json
d=json.load(open('gpurun_out/r3.out')); print('RESULT', d['gpu_ms']['ms_integrate'], d['gpu_ms']['ms_total'], d['parity']['max_rel_err_above_1e-6_of_max'])"; }
run3 SONG_LOS_TC_SHAPE=2,4,8
run3 SONG_LOS_TC_SHAPE=4,3,7
run3 SONG_LOS_TC_SHAPE=4,3,8
run3 SONG_LOS_TC_SHAPE=4,2,10
