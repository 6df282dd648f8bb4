#!/bin/bash
# Developer A/B harness for the secondary workloads: bash scripts/ab_w.sh <workload> [steps] — times bench.py --workload for
# the in-tree build and every _ab/*.so back to back on the same GPU box and prints ms/pass plus the result counts.
W=${1:-c4}; STEPS=${2:-3}
for lib in bourse_b200/libbourse_b200.so _ab/*.so; do
  [ -f "$lib" ] || continue
  BOURSE_B200_LIB=$PWD/$lib python bench.py --workload $W --steps $STEPS --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$lib', '$W', 'ms/pass %.2f' % d['ms_per_step'], d['orders_per_pass'], d['trades_per_pass'])"
done
