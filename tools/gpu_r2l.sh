set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload sn_t_10percent_lmax2000 --steps 5 --warmup 3 --extra none --no-cpu-baseline > gpurun_out/r2_bench_cfg1_k3.json 2> gpurun_out/r2_bench_cfg1_k3.err; tail -3 gpurun_out/r2_bench_cfg1_k3.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_cfg1_k3.json')); print(d['value'], d['e2e']['ms_per_step'], d['e2e']['ms_each_step'], d['k3_integral'])"
S=$(date +%s); timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_default_n2.json 2> gpurun_out/r2_bench_default_n2.err; echo "bench N=2 wall: $(( $(date +%s) - S )) s"; tail -5 gpurun_out/r2_bench_default_n2.err; cat gpurun_out/r2_bench_default_n2.json | cut -c1-6000
