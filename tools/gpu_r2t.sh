#!/bin/bash
# run T (reproducible builds): per-warp addresses opaque, helpers shared per chunk, binary kernel at 7 CTAs/SM, short f32 kernel at 5 CTAs/SM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2t
mkdir -p $O
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 600 python tools/dev_sweep.py --workload $w --ef $ef --steps 20 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-125
}
for v in "" opaque share; do run "$v" c2 128; run "$v" c3 128 --nq-list 2500,1250,1; done
run f32s5 c2 128 --sweep "ring_bytes=8192,6144"
run opaque c4s 200
run bin7 c4s 200
