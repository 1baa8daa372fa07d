python -m pytest tests/test_apply_gpu.py -x -q -m gpu 2>&1 | tail -3
for cfg in "" "--sweep-r 4" "--sweep-deep 0" "--p 2" "--p 6" "--p 6 --sweep-r 2"; do
  echo "== $cfg"; python bench.py --no-cpu --no-trace --steps 20 --warmup 5 $cfg 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline'].get('other_kernels_ms'))
    elif 'rror' in l: print(l.strip())
"
done
