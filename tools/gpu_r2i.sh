#!/bin/bash
# run I: compute-sanitizer over the rewritten heap merges / queue window / prefetches
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2i
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_run.py > $O/memcheck.log 2>&1; echo "rc=$?" >> $O/memcheck.log
tail -4 $O/memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_run.py > $O/racecheck.log 2>&1; echo "rc=$?" >> $O/racecheck.log
grep -c "hazard" $O/racecheck.log; tail -4 $O/racecheck.log
