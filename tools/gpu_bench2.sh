#!/bin/bash
O=gpurun_out/bench2; mkdir -p $O
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -5
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err ) 2> $O/time_n2.txt
tail -5 $O/bench_n2.err; tail -3 $O/time_n2.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench2/bench_n2.json').read().strip().splitlines()[-1])
def show(k,r): print(k, "value %.0f"%r['value'], "ms/step %.2f"%r['ms_per_step'], "e2e %.0f"%r['e2e']['value'], r['scaling'] if 'scaling' in r else '', r.get('parity_checked'))
show('c3',d)
for k in ('c4','c5'): show(k,d[k])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 0 --impl reference 2>/dev/null | cut -c1-300
