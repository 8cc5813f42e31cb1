for cfg in "" "FEGPU_COMPACT=0" "FEGPU_COMPACT=0 FEGPU_GATHER_BATCH=1" "FEGPU_COMPACT=0 FEGPU_GATHER_LPN=16" "FEGPU_GATHER_BATCH=1" "FEGPU_GATHER_LPN=16"; do
  echo "== $cfg"
  env $cfg python profiles/bench_configs.py c4 c3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:12], 'fresh %.2f cached %.2f integ %.2f sym %.2f num %.2f'%(d['fresh_ms'],d['cached_ms'],d['integrate_ms'],d['symbolic_ms'],d['numeric_ms']))
"
done
