#!/bin/bash
# Same-box A/B of bench.py: alternates arms for <rounds> rounds and prints ms per sup+unsup pair.
#   scripts/ab.sh <rounds> <arm> [<arm> ...]     arm = "base" (the _ab/base worktree) or "-" (working tree as is) or
#                                                 "VAR=VAL[,VAR2=VAL2]" (working tree with that environment)
R=${1:-2}; shift
one() { ( cd "$1" && env $2 timeout 120 python bench.py --no-cpu-baseline --steps 300 2>/dev/null | python -c 'import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("%.4f ms  e2e %.4f  top %s %.1fus" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel"], 1e3*d["roofline"]["ms_per_launch"]))' ); }
for i in $(seq $R); do
  for arm in "$@"; do
    case "$arm" in
      base) echo "base: $(one _ab/base AB_DUMMY=1)";;
      -) echo "new: $(one . AB_DUMMY=1)";;
      *) echo "new[$arm]: $(one . "${arm//,/ }")";;
    esac
  done
done
