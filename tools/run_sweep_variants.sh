python -m pytest tests/test_apply_gpu.py -x -q -m gpu 2>&1 | tail -5
for cfg in "--sweep-deep 1 --sweep-r 4" "--sweep-deep 1 --sweep-r 2" "--sweep-deep 0 --sweep-r 4" "--sweep-deep 0 --sweep-r 2" "--sweep-deep 1 --sweep-r 2 --sweep-ncs 1" "--sweep-deep 1 --sweep-r 2 --sweep-ncs 3" "--sweep-deep 1 --p 2" "--sweep-deep 1 --p 2 --sweep-r 2" "--sweep-deep 1 --p 6" "--sweep-deep 1 --p 6 --sweep-r 4"; do
  echo "== $cfg"; python bench.py --no-cpu --no-trace --steps 20 --warmup 5 $cfg 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])
    elif 'rror' in l: print(l.strip())
"
done
