for l in 8 16 32; do echo LPN $l; FEGPU_ROWS_LPN=$l python profiles/prof_diffusion.py 256 | tail -1; done
for l in 8 16 32; do echo T10 LPN $l; FEGPU_ROWS_LPN=$l python profiles/bench_configs.py c3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sym %.2f num %.2f'%(d['symbolic_ms'],d['numeric_ms']))"; done
