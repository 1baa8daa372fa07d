#!/bin/bash
# k_sweep variants on config 4 (one B200): kernel time and fraction of the measured HBM bandwidth per variant.
# A library built with other -D switches can be compared through HSBP_LIB=/path/to/libhsbp_variant.so.
for cfg in "" "--sweep-r 4" "--sweep-deep 0" "--sweep-deep 0 --sweep-r 2" "--sweep-ncs 1" "--sweep-ncs 3" "--p 2" "--p 6" "--p 6 --sweep-p6-regs 128"; do
  echo "== ${cfg:-default}"; python bench.py --no-cpu --no-trace --steps 20 --warmup 5 $cfg 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline'].get('other_kernels_ms'))
    elif 'rror' in l: print(l.strip())
"
done
