#!/bin/bash
# run L: config 4 at full size — phase shares and an ncu capture of the binary kernel at ef = 800
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2l
mkdir -p $O
HB_LIB_VARIANT=phases timeout 600 python tools/dev_sweep.py --workload c4 --ef 800 --steps 3 --nq 20000 --parity 64 > $O/c4_phases.log 2>&1
grep -h '^{' $O/c4_phases.log | cut -c1-700
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 4 -c 1 -f -o $O/c4_ncu python tools/dev_sweep.py --workload c4 --ef 800 --steps 1 --nq 20000 --parity 64 > $O/c4_ncu.log 2>&1
grep -h '^{' $O/c4_ncu.log | cut -c1-200
ls -la $O
