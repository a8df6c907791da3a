#!/bin/bash
run() { echo -n "$1: "; shift; env "$@" python bench.py --workload c4 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"; }
run "batch 4Mi" A=1
run "batch 8Mi" AMOFB_BATCH_ATOMS=8388608
run "batch 16Mi" AMOFB_BATCH_ATOMS=16777216
