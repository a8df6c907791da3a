#!/bin/bash
run() { echo -n "$1: "; shift; env "$@" python bench.py --workload c4 --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']))"; }
for rep in 1 2; do
run "prefetch 32" A=1
run "no prefetch" AMOFB_LIB=experiments/build/libamofb_nopf.so
run "prefetch 8" AMOFB_LIB=experiments/build/libamofb_pf8.so
run "prefetch 64" AMOFB_LIB=experiments/build/libamofb_pf64.so
done
