#!/bin/bash
timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python bench.py --workload c4 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c4', round(d['value']), round(d['e2e']['value']))"
