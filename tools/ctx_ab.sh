#!/bin/bash
# A/B of the context-coded initial keys (JP_BWT_FWD_CTXKEYS, JP_BWT_FWD_KEYPASSES) on the bench inputs and on real text.
cd "$(dirname "$0")/.."
for kind in markov2 uniform; do
  JP_BWT_FWD_CTXKEYS=0 timeout 120 python tools/fwd_ab.py $kind 64 2>&1 | tail -1
  for p in 5 6 7 8; do JP_BWT_FWD_KEYPASSES=$p timeout 120 python tools/fwd_ab.py $kind 64 2>&1 | tail -1; done
done
JP_BWT_TRACE_ROUNDS=1 JP_BWT_FWD_KEYPASSES=6 timeout 120 python tools/fwd_ab.py markov2 64 2>&1 | grep "jp_bwt keys" | tail -1
JP_BWT_FWD_CTXKEYS=0 timeout 120 python tools/fwd_ab.py markov2 256 2>&1 | tail -1
timeout 120 python tools/fwd_ab.py markov2 256 2>&1 | tail -1
for e in "JP_BWT_FWD_CTXKEYS=0" "JP_BWT_FWD_KEYPASSES=6" "JP_BWT_FWD_KEYPASSES=7" "JP_BWT_FWD_KEYPASSES=8"; do
  env $e timeout 300 python tools/real_text.py 64 2>&1 | tail -2
done
