for c in 86 0; do
  SWR_TC_CARVEOUT=$c timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('carveout',$c,'ms_per_step',round(d['ms_per_step'],4),[ (o['op'][:24],o['ms']) for o in d['ops_ms'][:9]])"
done
