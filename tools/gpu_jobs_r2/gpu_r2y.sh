#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_bad.py tests/test_gpu_classes.py tests/test_gpu_pair.py tests/test_gpu_guard.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2; do
python bench.py --workload c4 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gather', d['value'], d['e2e'])"
AMOFB_NO_HOST_GATHER=1 python bench.py --workload c4 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no gather', d['value'], d['e2e'])"
done
AMOFB_HOST_THREADS=8 python bench.py --workload c4 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gather 8 threads', d['value'], d['e2e'])"
AMOFB_HOST_THREADS=32 python bench.py --workload c4 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('gather 32 threads', d['value'], d['e2e'])"
nproc
