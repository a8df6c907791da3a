#!/bin/bash
O=gpurun_out/bench1; mkdir -p $O
( time python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err ) 2> $O/time_n1.txt
tail -3 $O/bench_n1.err; cat $O/time_n1.txt | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench1/bench_n1.json'))
def show(k,r): print(k, "value %.0f"%r['value'], "ms/step %.2f"%r['ms_per_step'], "e2e %.0f"%r['e2e']['value'], "roof", r['roofline']['bound'], "%.4f"%r['roofline']['frac'], "cpu", r.get('cpu_baseline',{}).get('value'), r.get('parity_checked'))
show('c2',d)
for k in ('c3','c4','c5'): show(k,d[k])
PY
