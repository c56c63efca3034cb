set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_k3_integral_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 1500 python tests/big/ab_knobs.py sn_pol_5percent_lmax3000 "SONG_LOS_TC_RMAX=2;SONG_LOS_TC_RMAX=4;SONG_LOS_TC_SHAPE=12/1/7;SONG_LOS_TC_SHAPE=12/1/7,SONG_LOS_TC_RMAX=3;SONG_LOS_TC_KCS=3;SONG_LOS_TC_KCS=4;SONG_LOS_TC_SHAPE=4/2/8;SONG_LOS_TC_SHAPE=4/2/10;SONG_LOS_TC_NP1=1;SONG_LOS_TC_SHAPE=8/1/8" 2>&1 | grep -E "RESULT|pairs,|Error" | tee gpurun_out/r2_ab_cfg3.log
