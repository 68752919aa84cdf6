#!/bin/bash
# compute-sanitizer over whole prim_run_subcycle_c calls of the CUDA dycore at the benchmarked dimensions
# (nlev 72, qsize 40) on a small mesh: memcheck, racecheck (shared-memory hazards inside the kernels), synccheck and
# initcheck. Summaries go to gpurun_out/sanitize_<tool>.txt; exits non-zero if any tool reports an error.
#   gpurun --timeout 900 -- 'bash scripts/sanitize.sh'            all four tools, ne4
#   NE=8 bash scripts/sanitize.sh memcheck racecheck              chosen tools, ne8
# profiles/r2_racecheck.txt is the round-2 record of the first two.
tools=("$@"); [ ${#tools[@]} -eq 0 ] && tools=(memcheck racecheck synccheck initcheck)
out=gpurun_out; mkdir -p $out
rc=0
for t in "${tools[@]}"; do
  timeout ${SANITIZE_TIMEOUT:-600} compute-sanitizer --tool $t --error-exitcode 9 \
    python scripts/prof_step.py --ne ${NE:-4} --qsize 40 --steps ${STEPS:-2} > $out/sanitize_$t.log 2>&1
  r=$?
  grep -E "SUMMARY|ERROR|Hazard|Error" $out/sanitize_$t.log | tail -5 > $out/sanitize_$t.txt
  echo "== $t: exit $r"; cat $out/sanitize_$t.txt
  [ $r -ne 0 ] && rc=1
done
exit $rc
