#!/bin/bash
# run D: tests on the new product library, short-row gather ceilings, ncu captures of the C2 / C4s kernels, phase shares
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2d
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
timeout 300 tools/gather4_bench > $O/gather4_bench.json 2> $O/gather4_bench.err; tail -c 300 $O/gather4_bench.json
for w in c2:128 c4s:200; do
  n=${w%%:*}; ef=${w##*:}
  timeout 300 python tools/dev_sweep.py --workload $n --ef $ef --steps 10 --device-build > $O/${n}_prod.log 2>&1
  HB_LIB_VARIANT=phases timeout 300 python tools/dev_sweep.py --workload $n --ef $ef --steps 10 --device-build > $O/${n}_phases.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 4 -c 1 -f -o $O/${n}_ncu python tools/dev_sweep.py --workload $n --ef $ef --steps 2 --device-build > $O/${n}_ncu.log 2>&1
done
grep -h '^{' $O/c2_prod.log $O/c2_phases.log $O/c4s_prod.log $O/c4s_phases.log | cut -c1-600
ls -la $O
