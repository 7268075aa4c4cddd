#!/bin/bash
# run B: merge block size x occupancy A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2b
mkdir -p $O
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 300 python tools/dev_sweep.py --workload $w --ef $ef --steps 10 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-330
}
for v in "" mb2 mb4; do run "$v" c3 128 --nq-list 1250,1; done
for v in "" mb2 mb4; do run "$v" c2 128; done
for v in "" bin5 bin6 bin7 bin8 mb2 mb4 mb2bin6 mb2bin8 mb4bin6 mb4bin8; do run "$v" c4s 200; done
