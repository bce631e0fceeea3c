#!/bin/bash
# Same-box A/B of bench.py: alternates arms for <rounds> rounds and prints ms per sup+unsup pair.
#   scripts/ab.sh <rounds> <arm> [<arm> ...]     arm = "base" (the _ab/base worktree) or "-" (working tree as is) or
#                                                 "VAR=VAL[,VAR2=VAL2]" (working tree with that environment)
# The base arm needs a built copy of the commit to compare against (its .so travels to the GPU box with the snapshot):
#   git worktree add -f _ab/base <commit> && (cd _ab/base && python __graft_entry__.py build)     # _ab/ is git-ignored
# Run it in ONE gpurun call: boxes differ by a few per cent, runs on the same box repeat to ~0.1 %.
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
