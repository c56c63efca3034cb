# 2 GPUs: multi-device drop-in tests + strong-scaling bench on configs[1]
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 600 python -m pytest tests/test_dropin.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload sn_t_10percent_lmax2000 --steps 10 --warmup 3 --extra none > gpurun_out/r2_bench_cfg1_n2.json 2> gpurun_out/r2_bench_cfg1_n2.err; tail -20 gpurun_out/r2_bench_cfg1_n2.err; cat gpurun_out/r2_bench_cfg1_n2.json
SONG_LOS_DEVICES=2 timeout 600 python tests/big/time_dropin.py --reps 4 --verbose 3 --oracle-pairs 4 2>&1 | tail -40
