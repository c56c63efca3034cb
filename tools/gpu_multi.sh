# N-GPU bench line under torch.distributed.run (the way the driver launches it)
set -x
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1; lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" >> gpurun_out/topo_n$N.txt; free -g >> gpurun_out/topo_n$N.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
if [ -z "$SKIP_REF" ]; then timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.json 2>> gpurun_out/bench_n$N.err; cat gpurun_out/bench_ref_n$N.json | cut -c1-300; fi
