#!/bin/bash
# run P: per-warp addresses made opaque (no rematerialisation) — is the kernel now insensitive to small edits? keep / dedupe / cold A/B again
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2p
mkdir -p $O
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 600 python tools/dev_sweep.py --workload $w --ef $ef --steps 20 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-110
}
for v in "" keep dedupef32 cold; do run "$v" c2 128; run "$v" c3 128 --nq-list 1250,1; done
run "" c4s 200
