# laps of the transfer2_init leg of bench.py under torchrun at N=2 (2440 ms there against 1488 ms in a plain process)
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --e2e-steps 2 --extra none --no-cpu-baseline --no-k3-integral --dropin-verbose 3 > gpurun_out/r2r_bench_n2_verbose.log 2>&1
grep -v "^{" gpurun_out/r2r_bench_n2_verbose.log | tail -40
