# ramped batches + early block-wise D2H in the drop-in (N=1, N=2), single-slot band shards at N=2
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
SONG_LOS_DEVICES=1 timeout 900 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 3 --verbose 3 --oracle-pairs 4 > gpurun_out/r2t_dropin_cfg3_n1.log 2>&1
SONG_LOS_DEVICES=2 timeout 900 python tests/big/time_dropin.py sn_pol_5percent_lmax3000 --reps 3 --verbose 3 --oracle-pairs 4 > gpurun_out/r2t_dropin_cfg3_n2.log 2>&1
grep "transfer2_init #\|max rel" gpurun_out/r2t_dropin_cfg3_*.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 0 --extra none --no-cpu-baseline --no-k3-integral --bands 2 > gpurun_out/r2t_n2_bands2.json 2> gpurun_out/r2t_n2_bands2.err; tail -3 gpurun_out/r2t_n2_bands2.err
timeout 1200 $TR bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 3 --extra none --no-cpu-baseline --no-k3-integral > gpurun_out/r2t_n2_default.json 2> gpurun_out/r2t_n2_default.err; tail -3 gpurun_out/r2t_n2_default.err
python - <<'PY'
import json
for f in ("gpurun_out/r2t_n2_bands2.json", "gpurun_out/r2t_n2_default.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["value"], d["ms_per_step"], d["per_rank_ms"], d["parity"]["ok"], d.get("nccl_gather", {}).get("parity_of_gathered_array"), d["roofline"]["step_breakdown_ms"], d["roofline"]["frac_per_rank"], d["roofline"]["batch_k3_values"], d["e2e"].get("ms_each_step"), d["e2e"].get("bessel2_init_s"))
PY
