#!/bin/bash
# ncu --set full captures of single conv_tc instantiations (one launch each) on tools/ncu_target.py.
#   tools/ncu_capture.sh <tag> <batch> <mode> "<BLOCK_N> <STAGES> <MODE> <XF> <RT> <HALO> <skip> <name>" ...
# <skip> = matching launches to skip (the warm-up pass launches each instantiation as often as the profiled pass).
# The .ncu-rep files stay on the GPU box (8 MB each; gpurun returns at most 64 MiB): what comes back is the raw-metric
# page of every capture (csv) and the SASS source page with stall samples (csv.gz).
tag=$1; batch=$2; mode=$3; shift 3
mkdir -p gpurun_out /tmp/ncu_reps
for spec in "$@"; do
  set -- $spec
  pat="k_conv_tc<\\(int\\)$1, \\(int\\)$2, \\(int\\)$3, \\(bool\\)$4, \\(bool\\)$5, \\(bool\\)$6>"
  rep=/tmp/ncu_reps/${tag}_conv_tc_$8
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k "regex:$pat" -s $7 -c 1 -f -o $rep python tools/ncu_target.py $batch $mode > gpurun_out/ncu_${tag}_$8.log 2>&1
  echo "$8 rc=$? $(grep -c '==PROF==' gpurun_out/ncu_${tag}_$8.log) prof lines"
  if [ -f $rep.ncu-rep ]; then
    ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/${tag}_conv_tc_$8.raw.csv 2>/dev/null
    ncu -i $rep.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${tag}_conv_tc_$8.source.csv.gz
  fi
done
ls -la gpurun_out/${tag}_*.csv* 2>/dev/null
