#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench line, ncu launch list of the bench command, ncu --set full of the two hot kernels.
# Usage (from the repo root, on the B200 box): bash scripts/gpu_validate.sh <tag>
TAG=${1:-r1c}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee $O/pytest_gpu_$TAG.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -4 | tee $O/smoke_$TAG.log
timeout 600 python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; tail -2 $O/bench_$TAG.err
python - <<PY
import json
d = json.loads(open("$O/bench_$TAG.json").read().strip().splitlines()[-1])
print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["clocks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 1 --iters 100 --no-cpu-baseline > $O/ncu_bench_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cfr_iter|leaf_mlp' -s 6 -c 2 -f -o $O/prof_$TAG \
  python scripts/ncu_target.py > $O/ncu_full_$TAG.log 2>&1; tail -2 $O/ncu_full_$TAG.log
ls -la $O/prof_$TAG.ncu-rep
