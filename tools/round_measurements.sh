#!/bin/sh
# One-GPU measurement pass of a round: tests, bench lines, ncu launch list and full captures.
# Everything lands in gpurun_out/ (scratch); profiles/ gets the summaries afterwards.
R=${1:-r2}
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${R}_pytest_gpu.log
python bench.py > $O/${R}_bench_default.log 2> $O/${R}_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_reference.log 2> $O/${R}_bench_reference.err; echo "reference arm rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_tokenids50k.csv \
    python bench.py --steps 2 --warmup 1 --quick --secondary none > $O/${R}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:jaccard -s 6 -c 1 -f -o $O/${R}_jaccard_tokenids50k \
    python bench.py --steps 1 --warmup 1 --quick --secondary none > $O/${R}_ncu_full1.log 2>&1; echo "ncu tokenids rc=$?"
ncu --set full --clock-control none --import-source on -k regex:jaccard -s 2 -c 1 -f -o $O/${R}_jaccard_term200k \
    python bench.py --workload term200k --steps 1 --warmup 1 --quick --secondary none > $O/${R}_ncu_full2.log 2>&1; echo "ncu term rc=$?"
ncu --set full --clock-control none --import-source on -k regex:qratio_flat -s 2 -c 2 -f -o $O/${R}_qratio_flat_fuzzy20k \
    python bench.py --workload fuzzy20k --steps 1 --warmup 1 --quick --secondary none > $O/${R}_ncu_full3.log 2>&1; echo "ncu fuzzy rc=$?"
for w in term200k defs1m variable20k fuzzy20k fuzzyterm10k mesh50k; do
  python bench.py --workload $w --steps 2 --quick --secondary none > $O/${R}_bench_$w.log 2> $O/${R}_bench_$w.err; echo "bench $w rc=$?"
done
python tools/pack_timing.py > $O/${R}_pack_timing.log 2>&1
python tools/cold_compare.py > $O/${R}_cold.log 2>&1; echo "cold rc=$?"
python tools/bench_summary.py $O/${R}_bench_default.log $O/${R}_bench_term200k.log $O/${R}_bench_defs1m.log $O/${R}_bench_variable20k.log $O/${R}_bench_fuzzy20k.log $O/${R}_bench_fuzzyterm10k.log $O/${R}_bench_mesh50k.log | cut -c1-260
