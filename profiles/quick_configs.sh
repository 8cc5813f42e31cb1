#!/bin/bash
# usage: bash profiles/quick_configs.sh "c2 c4" [ENV=VAL ...]   -- one line per config: phase times of profiles/bench_configs.py
cfgs="$1"; shift
env "$@" python profiles/bench_configs.py $cfgs 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:14], 'fresh %.2f cached %.2f integ %.2f sym %.2f num %.2f'%(d['fresh_ms'],d['cached_ms'],d['integrate_ms'],d['symbolic_ms'],d['numeric_ms']))
"
