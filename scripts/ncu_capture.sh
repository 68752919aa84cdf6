#!/bin/bash
# Full ncu capture of the element kernels of the first dynamics+tracer step at ne30/q40 and of the
# remap, exported as CSV on the GPU box (the .ncu-rep files are too large to travel back).
#   scripts/ncu_capture.sh <tag> [source-kernel-regex ...]
tag=${1:-cap}; shift
out=gpurun_out
mkdir -p $out
rep=/tmp/${tag}
timeout 1200 ncu --set full --clock-control none --import-source on \
  --kernel-name regex:"${NCU_REGEX:-caar_kernel|hv_first|hv_second|hv_update|euler_qminmax|euler_hvpost|euler_advect|minmax_kernel|remap_kernel|dss_pair|dss_quad}" \
  ${NCU_SKIP:+--launch-skip $NCU_SKIP} \
  ${NCU_COUNT:+-c $NCU_COUNT} -f -o $rep python scripts/prof_step.py ${PROF_ARGS} > $out/${tag}.log 2>&1
tail -2 $out/${tag}.log
ncu -i $rep.ncu-rep --page raw --csv > $out/${tag}_raw.csv 2>/dev/null
for k in "$@"; do
  safe=$(echo "$k" | tr -c 'A-Za-z0-9_' '_')
  ncu -i $rep.ncu-rep --page source --csv --kernel-name regex:"$k" --launch-count 1 > $out/${tag}_src_${safe}.csv 2>/dev/null
done
ls -la $out | tail -20
