#!/bin/bash
# Developer A/B for the deep-book config: bash scripts/ab_c5.sh [pages_smem] — in-tree build and every _ab/*.so back to back
PS=${1:-192}
for lib in bourse_b200/libbourse_b200.so _ab/*.so; do
  [ -f "$lib" ] || continue
  BOURSE_B200_LIB=$PWD/$lib python bench.py --workload c5 --steps 2 --no-cpu --pages-smem $PS 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$lib', 'c5 pages_smem $PS', 'ms/pass %.1f' % d['ms_per_step'], d['orders_per_pass'], d['trades_per_pass'])"
done
