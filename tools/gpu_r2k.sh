set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for stride in 1 2 4 8 16; do
  timeout 600 python tests/big/ab_knobs.py sn_t_10percent_lmax2000 "SONG_LOS_TC_NO_ADAPT=1" --pair-stride $stride --reps 3 2>&1 | grep -E "RESULT|pairs,|Error" | tee -a gpurun_out/r2_ab_adapt.log
done
timeout 600 python tests/big/ab_knobs.py sn_pol_10percent_lmax2000 "SONG_LOS_TC_NO_ADAPT=1" --pair-stride 1 --reps 3 2>&1 | grep -E "RESULT|pairs,|Error" | tee -a gpurun_out/r2_ab_adapt.log
timeout 600 python tests/big/ab_knobs.py sn_pol_10percent_lmax2000 "SONG_LOS_TC_NO_ADAPT=1" --pair-stride 4 --reps 3 2>&1 | grep -E "RESULT|pairs,|Error" | tee -a gpurun_out/r2_ab_adapt.log
for stride in 2 4 8; do
  timeout 900 python tests/big/ab_knobs.py sn_pol_5percent_lmax3000 "SONG_LOS_TC_NO_ADAPT=1" --pair-stride $stride --reps 2 2>&1 | grep -E "RESULT|pairs,|Error" | tee -a gpurun_out/r2_ab_adapt.log
done
timeout 600 python tests/big/time_dropin.py --reps 5 --verbose 3 --oracle-pairs 4 2>&1 | tail -32
