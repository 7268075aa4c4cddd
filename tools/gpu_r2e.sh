#!/bin/bash
# run E: speculative visited prefetch A/B, ring loop cleanup, config 4 at full size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2e
mkdir -p $O
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 300 python tools/dev_sweep.py --workload $w --ef $ef --steps 10 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-200
}
for v in "" nospec; do run "$v" c3 128 --nq-list 1250,1; done
for v in "" nospec; do run "$v" c2 128; done
for v in "" nospec bin7; do run "$v" c4s 200; done
timeout 900 python tools/c4_full.py --out $O/c4_full.json > $O/c4_full.log 2>&1; tail -2 $O/c4_full.log | cut -c1-1500
