#!/bin/bash
# tools/chunk_states_sweep.sh : end-to-end-from-states legs of configs[2] against the state chunk size (QPPVM_CHUNK_STATES)
for c in ${1:-8192 16384 32768}; do
QPPVM_CHUNK_STATES=$c python bench.py --config 2 --steps 3 --warmup 3 --no-latency --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk_states $c value %.0f e2e_states %.0f pipelined %.0f rollout %.0f' % (d['value'], d['e2e_states']['value'], d['e2e_states']['pipelined_value'], d['rollout']['value']))"
done
