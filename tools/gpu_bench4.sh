#!/bin/bash
N=${1:-4}
O=gpurun_out/bench$N; mkdir -p $O
nvidia-smi -L | wc -l
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2> $O/time.txt
tail -4 $O/bench.err; grep real $O/time.txt
python - <<PY
import json
d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1])
def show(k,r): print(k, "value %.0f"%r['value'], "ms/step %.2f"%r['ms_per_step'], "e2e %.0f"%r['e2e']['value'], r.get('scaling',''), r.get('parity_checked'))
show('c3',d)
for k in ('c4','c5'): show(k,d[k])
PY
