#!/bin/bash
# same-box A/B of builds of libb200ir.so kept under image-retrieval-_b200/build/:
#   tests/tools/ab_libs.sh "libA.so libB.so ..." [bench args]      (two rounds over the list, step / kernel ms per run)
LIBS=$1; shift
D=image-retrieval-_b200
cp $D/libb200ir.so $D/build/lib_saved.so
for round in 1 2; do
  for v in $LIBS; do
    cp $D/build/$v $D/libb200ir.so
    python bench.py --no-cpu --no-side --no-strong --steps 20 --warmup 3 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$v', 'step %.3f ms' % d['ms_per_step'], 'e2e %.3f ms' % d['e2e']['ms_per_step'], {k: round(x,3) for k,x in r['kernels_ms'].items() if x > 0.05}, d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
cp $D/build/lib_saved.so $D/libb200ir.so
