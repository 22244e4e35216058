#!/bin/bash
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) "$@" 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -2; }
echo "== rmat22 N=8"; run 8 bench.py --gpus 8 --steps 2 --warmup 3 --workload rmat22 --no-e2e | tail -1 > gpurun_out/bench_n8_r22.json; python scratch/show.py gpurun_out/bench_n8_r22.json
echo "== rmat20mc N=8"; run 8 bench.py --gpus 8 --steps 2 --warmup 3 --workload rmat20mc --no-e2e | tail -1 > gpurun_out/bench_n8_mc.json; python scratch/show.py gpurun_out/bench_n8_mc.json
