#!/bin/bash
# round-2 session-2 run A: parity of the in-place merge + A/B timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out/r2a
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
timeout 600 python tools/dev_sweep.py --workload c3 --ef 128 --steps 20 --nq-list 5000,2500,1250,1 --out $O/c3.json > $O/c3.log 2>&1
timeout 300 python tools/dev_sweep.py --workload c2 --ef 128 --steps 20 --device-build --out $O/c2.json > $O/c2.log 2>&1
timeout 300 python tools/dev_sweep.py --workload c4s --ef 200 --steps 10 --device-build --out $O/c4s.json > $O/c4s.log 2>&1
for v in bin6 bin8 phases; do
  HB_LIB_VARIANT=$v timeout 300 python tools/dev_sweep.py --workload c4s --ef 200 --steps 10 --device-build --out $O/c4s_$v.json > $O/c4s_$v.log 2>&1
done
HB_LIB_VARIANT=phases timeout 300 python tools/dev_sweep.py --workload c2 --ef 128 --steps 10 --device-build --out $O/c2_phases.json > $O/c2_phases.log 2>&1
grep -h '^{' $O/*.log | cut -c1-400
