#!/bin/bash
# same-box A/B of resize builds: tests/tools/ab_resize.sh "libR0.so libR1.so" (libraries under image-retrieval-_b200/build/)
D=image-retrieval-_b200
cp $D/libb200ir.so $D/build/lib_saved.so
for round in 1 2; do for v in $1; do cp $D/build/$v $D/libb200ir.so; python bench.py --workload resize --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), 'images/s, kernel', round(d['roofline']['kernel_ms'], 4), 'ms')"; done; done
cp $D/build/lib_saved.so $D/libb200ir.so
