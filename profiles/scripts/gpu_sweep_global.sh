for c in 4 6 8 12; do
  for L in 256 1024; do
  S=500000; if [ $L = 1024 ]; then S=100000; fi
  TSFX_GLOBAL_CTAS=$c python bench.py --series $S --len $L --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ctas=$c L=$L', round(d['ms_per_step'],1), {k:round(v['ms'],1) for k,v in d['roofline']['groups'].items() if k in ('seq','peaks')})"
  done
done
