#!/bin/bash
# run J (2 GPUs): the driver's N=2 command (strong scaling line + config-5-shaped sharded record under the deadline guard), microbench again
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2j
mkdir -p $O
free -g | head -2 > $O/host.txt; nproc >> $O/host.txt
timeout 200 tools/gather4_bench > $O/gather4_bench.json 2> $O/gather4_bench.err
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu.json 2> $O/bench_2gpu.err ) 2> $O/bench_2gpu.time
tail -3 $O/bench_2gpu.time; grep "bench r0" $O/bench_2gpu.err | tail -20; tail -c 1500 $O/bench_2gpu.json
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q > $O/multigpu_tests.log 2>&1; tail -2 $O/multigpu_tests.log
cat $O/host.txt
