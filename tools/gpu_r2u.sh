#!/bin/bash
# run U (reproducible builds): lean bookkeeping (counters only on request, hoisted uniform tests) A/B + parity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2u
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_query_options.py tests/test_gpu_round2.py -m gpu -q 2>&1 | tail -2
run() { # variant workload ef extra...
  local v=$1 w=$2 ef=$3; shift 3
  HB_LIB_VARIANT=$v timeout 600 python tools/dev_sweep.py --workload $w --ef $ef --steps 20 --device-build "$@" > $O/${w}_${v:-prod}.log 2>&1
  echo "== $w ${v:-prod}"; grep -h '^{' $O/${w}_${v:-prod}.log | cut -c1-125
}
for v in "" nolean; do run "$v" c2 128; run "$v" c3 128 --nq-list 1250,1; run "$v" c4s 200; done
