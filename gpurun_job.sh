mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 10 2>&1 | tail -2 | tee gpurun_out/bench_n2.json | cut -c1-700
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference 2>&1 | tail -1 | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload c5 --rollout --steps 3 --warmup 2 2>&1 | tail -1 | tee gpurun_out/bench_n2_c5_rollout.json | cut -c1-900
