#!/bin/bash
# value / e2e of the inverse with several blocks in flight, for walker residency and caller counts
cd "$(dirname "$0")/.."
for cfg in "5 4" "4 4" "3 4" "4 6" "3 6" "3 8" "2 8"; do
  set -- $cfg
  JP_BWT_INV_WBLOCKS_PER_SM=$1 JP_BWT_MAX_CTX=$2 timeout 200 python bench.py --steps 10 --warmup 3 --no-forward --no-cpu-baseline --no-configs --callers $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('wblocks/SM=$1 callers=$2 value', d['value'], 'e2e', d['e2e']['value'], 'single', d['single_stream']['value'], d['parity']['round_trip'])"
done
