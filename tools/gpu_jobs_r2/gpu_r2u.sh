#!/bin/bash
mkdir -p gpurun_out/r2u
timeout 600 python -m pytest tests/test_gpu_bad.py tests/test_gpu_classes.py tests/test_gpu_guard.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
python bench.py --workload c4 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lists', d['value'], d['e2e']['value'])"
AMOFB_BAD_ONE_LIST=1 python bench.py --workload c4 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('one list', d['value'], d['e2e']['value'])"
AMOFB_BAD_CELL_WIDEN=1.25 python bench.py --workload c4 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('lists widen 1.25', d['value'], d['e2e']['value'])"
done
