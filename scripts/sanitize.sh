#!/bin/bash
# compute-sanitizer over short runs of every kernel family; logs under gpurun_out/.
# NOTE (round 2): compute-sanitizer is closed on the GPU pool this repository is developed on (the client refuses to start it:
# profiles/sanitize_r02.txt); the script is kept for boxes where it is available.  The substitute checks are numerical: see
# tests/test_gpu_next_rows.py::test_initialize_with_every_slot_in_use_matches_oracle and the watchdog traps of the mbarrier waits.
#   scripts/sanitize.sh [memcheck|racecheck|synccheck ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in "${@:-memcheck racecheck}"; do :; done
TOOLS="${*:-memcheck racecheck}"
for tool in $TOOLS; do
  for fam in persistent tile split bigr rls64 aux; do
    log=gpurun_out/sanitize_${tool}_${fam}.txt
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_run.py $fam > $log 2>&1
    echo "== $tool $fam: exit $? | $(grep -c SANITIZE_RAN $log) ran | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
  done
done
