#!/bin/bash
# Where the wall time of the reference CLI + our stage goes at N GPUs: start-up of the contexts, the run, the exit.
#   bash tools/pipeline_phases.sh <corpus blocks> <devices, e.g. 0,1>
set -u
cd "$(dirname "$0")/.."
blocks=${1:-16}; devs=${2:-0}
d=$(mktemp -d)
python - <<PY
import sys; sys.path.insert(0, ".")
import synth
with open("$d/corpus.bin", "wb") as f:
    for b in range($blocks): synth.gen("markov2", 64 << 20, 100 + b).tofile(f)
PY
for dv in 0 "$devs"; do
  echo "== JP_BWT_DEVICES=$dv compress"
  ( time JP_BWT_DEVICES=$dv JP_BWT_TRACE=1 oracle/_ref/Jampack_shim c $d/corpus.bin $d/out.jam -b64 -t16 > /dev/null ) 2>&1 | grep -E "warmup|trace\]|real"
  echo "== JP_BWT_DEVICES=$dv decompress"
  ( time JP_BWT_DEVICES=$dv JP_BWT_TRACE=1 oracle/_ref/Jampack_shim d $d/out.jam $d/back -t16 > /dev/null ) 2>&1 | grep -E "warmup|trace\]|real"
done
echo "== reference (CPU stage)"
( time oracle/_ref/Jampack_ref c $d/corpus.bin $d/ref.jam -b64 -t16 > /dev/null ) 2>&1 | grep real
( time oracle/_ref/Jampack_ref d $d/ref.jam $d/back -t16 > /dev/null ) 2>&1 | grep real
cmp $d/out.jam $d/ref.jam && echo "jam identical"
rm -rf $d
