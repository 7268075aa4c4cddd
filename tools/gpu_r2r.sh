#!/bin/bash
# run R: the round's main A/Bs repeated with reproducible builds (speculative visited prefetch, short-row kernel and ring, binary occupancy, merge block)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2r
mkdir -p $O
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 600 python tools/dev_sweep.py --workload $w --ef $ef --steps 20 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-130
}
run "" c2 128 --sweep "ring_short=1,0;ring_bytes=8192,12288"
run nospec c2 128
run g4 c2 128
run g8 c2 128
run "" c3 128 --nq-list 1250,1
run nospec c3 128 --nq-list 1250,1
run g4 c3 128 --nq-list 1250,1
run "" c4s 200
run nospec c4s 200
run bin4 c4s 200
run g4 c4s 200
