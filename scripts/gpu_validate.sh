#!/bin/bash
# One GPU-box pass for the tracked evidence of a build: parity tests, smoke, the bench lines, the ncu launch list of the bench command,
# ncu --set full of the two hot kernels inside a self-play wave, the clock64 timeline of the value-net kernel.
# Usage (from the repo root, on the B200 box): bash scripts/gpu_validate.sh <tag>         (then, in the build container:
#                                              python scripts/make_profile_summary.py <tag>)
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
rm -f $O/parity_notes.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee $O/${TAG}_pytest_gpu.log
cp $O/parity_notes.log $O/${TAG}_parity_notes.log 2>/dev/null
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -4 | tee $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -2 $O/${TAG}_bench.err
timeout 600 python bench.py --workload solve --steps 10 --warmup 3 > $O/${TAG}_bench_solve.json 2> $O/${TAG}_bench_solve.err; tail -2 $O/${TAG}_bench_solve.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; tail -2 $O/${TAG}_bench_reference.err
for w in config4 config5; do timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; tail -1 $O/${TAG}_bench_$w.err; done
timeout 120 python scripts/tc_trace.py > $O/${TAG}_tc_trace.log 2>&1
for net in tcx2 zero; do timeout 200 python scripts/datagen_probe.py --net $net; done > $O/${TAG}_datagen_probe.log 2>&1
timeout 120 python scripts/gelu_table.py 2>&1 | grep -v Warn > $O/${TAG}_gelu_table.log
timeout 120 python scripts/cfr_probe.py > $O/${TAG}_cfr_probe.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --iters 64 --no-cpu-baseline > $O/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cfr_iter_d2|leaf_mlp_tc3' -s 600 -c 2 -f -o $O/${TAG}_prof \
  python scripts/datagen_probe.py --iters 64 --waves 1 --warm 6 > $O/${TAG}_ncu_full.log 2>&1; tail -2 $O/${TAG}_ncu_full.log
ls -la $O | tail -20
