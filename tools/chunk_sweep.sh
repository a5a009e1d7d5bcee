#!/bin/bash
# tools/chunk_sweep.sh : end-to-end legs of configs[2] against the host-path chunk size (QPPVM_CHUNK records, 4x that many states)
for c in ${1:-1024 2048 4096 8192}; do
QPPVM_CHUNK=$c python bench.py --config 2 --steps 3 --warmup 3 --no-latency --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c value %.0f e2e_records %.0f pipelined %.0f e2e_states %.0f pipelined %.0f' % (d['value'], d['e2e_records']['value'], d['e2e_pipelined']['value'], d['e2e_states']['value'], d['e2e_states']['pipelined_value']))"
done
