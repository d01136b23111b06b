for f in 8 16 32 64; do
  SWR_LAZY_FLUSH=$f timeout 300 python bench.py --steps 256 --warmup 10 --no-cpu-baseline --no-saturated --no-scopes 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('flush_every',$f,'ms_per_step',round(d['ms_per_step'],4),'opt',round(d['roofline_optimizer']['kernel_ms_per_step'],4),[ (o['op'][:20],o['ms']) for o in d['ops_ms'] if 'adam' in o['op']])"
done
SWR_LAZY_ADAM=0 timeout 300 python bench.py --steps 256 --warmup 10 --no-cpu-baseline --no-saturated --no-scopes 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('dense ms_per_step',round(d['ms_per_step'],4))"
