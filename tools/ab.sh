#!/bin/sh
# A/B of kernel build variants on one box: tools/ab.sh <workload> <lib>...   ("product" = the in-tree library)
w=$1; shift
for lib in "$@"; do
  if [ "$lib" = product ]; then unset NSM_B200_LIB; else export NSM_B200_LIB=$PWD/$lib; fi
  python bench.py --workload $w --steps 3 --quick > gpurun_out/ab_$(basename $lib .so)_$w.log 2> gpurun_out/ab_$(basename $lib .so)_$w.err
  echo "$lib rc=$?"; python tools/bench_summary.py gpurun_out/ab_$(basename $lib .so)_$w.log | cut -c1-150
done
