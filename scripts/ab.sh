#!/bin/bash
# Developer A/B harness: times the C3 k_sim pass for every library variant in _ab/*.so plus the in-tree build,
# back to back on the same GPU box.  Usage (under gpurun): bash scripts/ab.sh [steps] [extra bench.py args]
STEPS=${1:-5}; shift
for lib in bourse_b200/libbourse_b200.so _ab/*.so; do
  [ -f "$lib" ] || continue
  BOURSE_B200_LIB=$PWD/$lib python bench.py --steps $STEPS --warmup 3 --no-cpu "$@" 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('$lib', '$*', 'ms/pass %.2f' % d['ms_per_step'], 'orders/s %.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'], 'clk', d['clocks']['sm_mhz'], 'chk', d['l1_checksums'][0] % 1000003)
    elif 'rror' in ln: print(ln.strip())
"
done
