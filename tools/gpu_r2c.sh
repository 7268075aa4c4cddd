#!/bin/bash
# run C: guarded merge blocks, f32 ring kernel at higher occupancy with smaller rings (C2), binary kernel occupancy
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c
mkdir -p $O
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 300 python tools/dev_sweep.py --workload $w --ef $ef --steps 10 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-200
}
for v in "" g2 g3; do run "$v" c3 128 --nq-list 1250,1; done
for v in "" g2 f32b4 f32b5 g2f32b4; do run "$v" c2 128 --sweep "ring_bytes=12288,8192,6144,4096"; done
for v in "" g2 bin5 bin6 bin7 g2bin5 g2bin6; do run "$v" c4s 200; done
