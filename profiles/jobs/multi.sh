# multi-GPU bench lines: bash profiles/jobs/multi.sh N
N=$1; mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$T bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_n${N}_reference.json | cut -c1-400
$T bench.py --gpus $N --steps 300 --warmup 10 2>&1 | tail -1 > gpurun_out/bench_n${N}.json; cut -c1-300 gpurun_out/bench_n${N}.json
$T bench.py --gpus $N --workload c5 --rollout --steps 3 --warmup 1 --no-extra --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_n${N}_c5_rollout.json; cut -c1-300 gpurun_out/bench_n${N}_c5_rollout.json
