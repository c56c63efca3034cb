timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for d in ${DBGS:-0 4 5 7}; do
  echo "== dbg $d"; SONG_LOS_TC_DBG=$d timeout 300 python tools/time_configs.py --mode fast --reps 2 --oracle-pairs 20 ${CONFIGS:-sn_t_10percent_lmax2000} 2>&1 | grep -E "^FAST" | tail -2 | grep -o "'ms_integrate': [0-9.]*\|max rel err above floor [0-9.e+-]*"
done
