for cfg in "1" "2" "3" "4"; do
  TSFX_STREAMS=$cfg python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('streams=$cfg', round(d['ms_per_step'],1), 'ms/step', round(d['value']))"
done
