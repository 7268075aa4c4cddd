#!/bin/bash
# final evidence of round 2 (one B200): GPU tests, the driver's bench command, ncu launch list + full captures (C3, C2, C4s), config 4 at full size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2final
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/bench.time
tail -3 $O/bench.time; tail -c 300 $O/bench.json; echo
timeout 300 python bench.py --steps 20 --warmup 5 --ef 128 --extras 0 > $O/bench_ef128.json 2> $O/bench_ef128.err; tail -c 200 $O/bench_ef128.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --ef 128 --extras 0 > $O/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 6 -c 1 -f -o $O/c3_ncu python bench.py --steps 2 --warmup 3 --ef 128 --extras 0 > $O/c3_ncu.log 2>&1
for w in c2:128 c4s:200; do
  n=${w%%:*}; ef=${w##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 4 -c 1 -f -o $O/${n}_ncu python tools/dev_sweep.py --workload $n --ef $ef --steps 2 --device-build > $O/${n}_ncu.log 2>&1
done
timeout 900 python tools/c4_full.py --out $O/c4_full.json > $O/c4_full.log 2>&1; tail -1 $O/c4_full.log | cut -c1-400
ls -la $O
