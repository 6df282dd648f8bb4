for e in 148 592 1184 2368 3552 4096 8192; do
python bench.py --steps 3 --warmup 3 --no-cpu --envs $e 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('envs', $e, 'ms/pass %.2f' % d['ms_per_step'], 'orders/s %.3e' % d['value'])
    elif 'rror' in ln: print(ln.strip())
"
done
