#!/bin/bash
# same-box A/B: alternate the baseline worktree (_ab/base) and the working tree; prints ms per sup+unsup pair
# usage: scripts/ab.sh <rounds> [VAR=VALUE ...]   (extra env assignments apply to a third arm "new+env")
R=${1:-2}; shift
EXTRA="$@"
one() { ( cd "$1" && env $2 timeout 120 python bench.py --no-cpu-baseline --steps 300 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.4f ms  e2e %.4f  top %s %.1fus" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel"], 1e3*d["roofline"]["ms_per_launch"]))' ); }
for i in $(seq $R); do
  echo "base   : $(one _ab/base GCCVAE_GATE_BWD_STREAM=main)"
  echo "new    : $(one . X=1)"
  if [ -n "$EXTRA" ]; then echo "new+env: $(one . "$EXTRA")"; fi
done
