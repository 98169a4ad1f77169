#!/bin/sh
# One-GPU measurement pass of a round: tests, bench lines, ncu launch list and full captures.
# Everything lands in gpurun_out/ (scratch); profiles/ gets the summaries afterwards.
R=${1:-r1}
O=gpurun_out
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_default.log 2> $O/bench_default.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches_tokenids50k.csv \
    python bench.py --steps 2 --warmup 1 --quick > $O/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:jaccard -s 3 -c 1 -f -o $O/${R}_jaccard_tokenids50k \
    python bench.py --steps 1 --warmup 1 --quick > $O/ncu_full1.log 2>&1; echo "ncu tokenids rc=$?"
ncu --set full --clock-control none --import-source on -k regex:jaccard -s 2 -c 1 -f -o $O/${R}_jaccard_term200k \
    python bench.py --workload term200k --steps 1 --warmup 1 --quick > $O/ncu_full2.log 2>&1; echo "ncu term rc=$?"
ncu --set full --clock-control none --import-source on -k regex:qratio -s 2 -c 2 -f -o $O/${R}_qratio_fuzzy20k \
    python bench.py --workload fuzzy20k --steps 1 --warmup 1 --quick > $O/ncu_full3.log 2>&1; echo "ncu fuzzy rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pack_sets -s 4 -c 2 -f -o $O/${R}_pack_term1m \
    python tools/pack_timing.py > $O/ncu_full4.log 2>&1; echo "ncu pack rc=$?"
for w in term200k defs1m term1m variable20k fuzzy20k fuzzyterm10k; do
  python bench.py --workload $w --steps 2 > $O/bench_$w.log 2> $O/bench_$w.err; echo "bench $w rc=$?"
done
python tools/pack_timing.py > $O/pack_timing.log 2>&1
python tools/bench_summary.py $O/bench_default.log $O/bench_term200k.log $O/bench_defs1m.log $O/bench_term1m.log $O/bench_variable20k.log $O/bench_fuzzy20k.log $O/bench_fuzzyterm10k.log
