for i in 1 2 3; do
  for ov in 1 0; do
    FEGPU_OVERLAP=$ov python bench.py --steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('overlap=$ov', 'ms_per_step %.2f'%d['ms_per_step'], 'e2e %.1fM'%(d['e2e']['value']/1e6), 'cached %.2f'%d['cached']['ms_per_step'], d['clocks'])"
  done
done
