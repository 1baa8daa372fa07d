for lib in "" hybridsbp_b200/variants/libhsbp_e2.so hybridsbp_b200/variants/libhsbp_e4.so hybridsbp_b200/variants/libhsbp_e5.so; do
  echo "== lib ${lib:-default}"; HSBP_LIB=${lib:+$PWD/$lib} python bench.py --no-cpu --no-trace --steps 20 --warmup 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline'].get('other_kernels_ms'))
    elif 'rror' in l: print(l.strip())
"
done
for cfg in "--p 6" "--p 6 --sweep-p6-regs 168"; do
  echo "== $cfg"; python bench.py --no-cpu --no-trace --steps 20 --warmup 5 $cfg 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline'].get('other_kernels_ms'))
    elif 'rror' in l: print(l.strip())
"
done
