#!/bin/bash
# run K (4 GPUs): the driver's N=4 command — strong scaling line + config-5-shaped sharded record (fused exchange between 4 peers)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2k
mkdir -p $O
free -g | head -2 > $O/host.txt; nproc >> $O/host.txt; nvidia-smi topo -m >> $O/host.txt 2>&1
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_4gpu.json 2> $O/bench_4gpu.err ) 2> $O/bench_4gpu.time
tail -3 $O/bench_4gpu.time; grep "bench r0" $O/bench_4gpu.err | tail -12; tail -c 1800 $O/bench_4gpu.json
head -3 $O/host.txt
