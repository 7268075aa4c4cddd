#!/bin/bash
# run N (8 GPUs): the driver's N=8 command — strong scaling line + config 5 at shape (8 shards x 6.25M x 768, 100k queries)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2n
mkdir -p $O
free -g | head -2 > $O/host.txt; nproc >> $O/host.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_8gpu.json 2> $O/bench_8gpu.err ) 2> $O/bench_8gpu.time
tail -3 $O/bench_8gpu.time; grep "bench r0" $O/bench_8gpu.err | tail -12; tail -c 1500 $O/bench_8gpu.json
head -3 $O/host.txt
