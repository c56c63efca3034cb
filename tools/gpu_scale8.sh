# 8 GPUs: configs[3] on 2 pair groups x 4 k3 bands (value, gather + parity, transfer2_init on all GPUs), then configs[4]
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv | head -3; nproc; free -g | head -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
S=$(date +%s); timeout 1100 $TR bench.py --gpus 8 --steps 3 --warmup 3 --e2e-steps 2 --extra none --no-cpu-baseline --dropin-verbose 2 > gpurun_out/r2_scale8_n8_cfg3.json 2> gpurun_out/r2_scale8_n8_cfg3.err; echo "wall: $(( $(date +%s) - S )) s"; tail -3 gpurun_out/r2_scale8_n8_cfg3.err
S=$(date +%s); timeout 900 $TR bench.py --gpus 8 --steps 2 --warmup 3 --e2e-steps 0 --extra none --no-cpu-baseline --workload sn_pol_5percent_lmax4000 > gpurun_out/r2_scale8_n8_cfg4.json 2> gpurun_out/r2_scale8_n8_cfg4.err; echo "wall: $(( $(date +%s) - S )) s"; tail -3 gpurun_out/r2_scale8_n8_cfg4.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_scale8_n8_cfg3.json", "gpurun_out/r2_scale8_n8_cfg4.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["value"], d["ms_per_step"], [round(x, 1) for x in d["per_rank_ms"]], d["parity"]["ok"], d.get("nccl_gather", {}).get("parity_of_gathered_array"), d["roofline"]["step_breakdown_ms"], [round(x, 3) for x in d["roofline"]["frac_per_rank"]], d["roofline"]["batch_k3_values"], d["e2e"].get("ms_each_step"), d["e2e"].get("bessel2_init_s"), d["config"]["parallelism"][:120])
PY
grep -v "^{" gpurun_out/r2_scale8_n8_cfg3.json | grep "GPU\|since\|pairs in" | head -60
