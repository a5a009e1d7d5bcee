#!/bin/bash
# tools/ab.sh "variants..." "configs..." : device-resident solves/s of each library variant ("main" = the in-tree build)
for v in $1; do
  lib=qppvm_b200/variants/libqppvm_b200_$v.so; [ "$v" = main ] && lib=qppvm_b200/libqppvm_b200.so
  for c in $2; do
    QPPVM_B200_LIB=$PWD/$lib python bench.py --config $c --steps 10 --warmup 3 --no-latency --no-cpu-baseline --fast 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v cfg$c value %.0f ms/step %.3f conv %.4f' % (d['value'], d['ms_per_step'], d['converged_frac']))"
  done
done
