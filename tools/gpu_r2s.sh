# k3-band schedule: GPU suite (slices against whole pairs), N=2 value with 2 bands against 2 pair groups, e2e at N=2 with the nested-OpenMP fix
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 0 --extra none --no-cpu-baseline --no-k3-integral --bands 2 > gpurun_out/r2s_n2_bands2.json 2> gpurun_out/r2s_n2_bands2.err; tail -c 2500 gpurun_out/r2s_n2_bands2.json | head -c 1200; tail -3 gpurun_out/r2s_n2_bands2.err
timeout 1200 $TR bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 2 --extra none --no-cpu-baseline --no-k3-integral > gpurun_out/r2s_n2_default.json 2> gpurun_out/r2s_n2_default.err; tail -3 gpurun_out/r2s_n2_default.err
python - <<'PY'
import json
for f in ("gpurun_out/r2s_n2_bands2.json", "gpurun_out/r2s_n2_default.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d["value"], d["ms_per_step"], d["per_rank_ms"], d["parity"], d.get("nccl_gather", {}).get("parity_of_gathered_array"), d["roofline"]["step_breakdown_ms"], d["roofline"]["frac_per_rank"], d["e2e"].get("ms_each_step"), d["e2e"].get("bessel2_init_s"))
PY
