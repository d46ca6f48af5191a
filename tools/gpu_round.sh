#!/bin/bash
# One call on the B200 box for the evidence a round commits: GPU tests, smoke, the bench line, ncu launch lists of the
# measured step and of the bench command, --set full captures of the inverse kernels and of one whole forward call
# (-> profiles/traffic.json), the configs table. Everything lands in gpurun_out/ (merged back by gpurun); copy what is
# to be judged into profiles/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tests] [bench] [ncu]'   (default: all)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what="${*:-tests bench ncu}"
R=${ROUND:-r02}
has() { [[ " $what " == *" $1 "* ]]; }

if has tests; then
	timeout 1000 python -m pytest tests -m gpu -x -q --timeout=400 2>&1 | tail -6 | tee gpurun_out/tests.log
	timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 | tee gpurun_out/smoke.log
fi
if has ncu; then
	timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
		-k regex:"k_inv_walk_stream|k_inv_rank_packed|k_inv_place|k_inv_clear_text|k_inv_lf|k_inv_hist" -o gpurun_out/inv_single_full \
		python tools/profile_step.py 64 markov2 inv > gpurun_out/ncu_full.log 2>&1
	ncu -i gpurun_out/inv_single_full.ncu-rep --page raw --csv > gpurun_out/inv_single_ncu_full_$R.csv 2>/dev/null
	ncu -i gpurun_out/inv_single_full.ncu-rep --page source --csv -k regex:k_inv_walk_stream > gpurun_out/inv_walk_stream_source_$R.csv 2>/dev/null
	ncu -i gpurun_out/inv_single_full.ncu-rep --page source --csv -k regex:k_inv_place > gpurun_out/inv_place_source_$R.csv 2>/dev/null
	rm -f gpurun_out/inv_single_full.ncu-rep
	python tools/ncu_traffic.py gpurun_out/inv_single_ncu_full_$R.csv --update profiles/traffic.json --key inverse_walk_single \
		--sum k_inv_walk_stream,k_inv_rank_packed,k_inv_clear_text,k_inv_place | tail -2 | tee gpurun_out/traffic.log
	# one whole forward call, every kernel of it
	timeout 400 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/fwd_full \
		python tools/profile_step.py 64 markov2 fwd > gpurun_out/ncu_fwd_full.log 2>&1
	ncu -i gpurun_out/fwd_full.ncu-rep --page raw --csv > gpurun_out/fwd_ncu_full_$R.csv 2>/dev/null
	rm -f gpurun_out/fwd_full.ncu-rep
	python tools/ncu_traffic.py gpurun_out/fwd_ncu_full_$R.csv --update profiles/traffic.json --key forward --sum-all | tail -1 | tee -a gpurun_out/traffic.log
	cp profiles/traffic.json gpurun_out/traffic.json
	timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step_$R.csv \
		python tools/profile_step.py 64 markov2 both > gpurun_out/ncu_step.log 2>&1
	timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_bench_$R.csv \
		python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/ncu_bench.log 2>&1
fi
if has bench; then
	timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
	python - <<PY
import json
d = json.load(open("gpurun_out/bench_$R.json")); f = d.get("forward", {})
print("inverse", d["value"], "single", d["single_stream"]["value"], "e2e", d["e2e"]["value"], "| forward", f.get("value"), "e2e", f.get("e2e", {}).get("value"),
      "| cpu", d.get("cpu_baseline", {}).get("value"), f.get("cpu_baseline", {}).get("value"), "| legacy cuda", d.get("legacy_cuda_baseline"), "|", d["parity"], d["clocks"]["reasons"])
PY
fi
ls -la gpurun_out | tail -12
