#!/bin/bash
# compute-sanitizer racecheck + memcheck of the fused kernel on a few instances of each problem family, both arithmetics
# (run on a GPU box)
set -x
for ARITH in 0 1; do
  export PMB_ARITH=$ARITH
  PMB_MAX_ITER=3 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/profile_step.py mobile_robot 4 2 2>&1 | grep -E "RACECHECK|Error|Warning" | head
  PMB_MAX_ITER=2 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/profile_step.py cstr 3 1 2>&1 | grep -E "RACECHECK|Error|Warning" | head
  PMB_MAX_ITER=2 PMB_BLOCK_BFGS=1 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/profile_step.py parking 3 1 2>&1 | grep -E "RACECHECK|Error|Warning" | head
  PMB_MAX_ITER=3 PMB_PRECOND=2 PMB_FILTER_LS=1 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/profile_step.py mobile_robot 3 2 2>&1 | grep -E "RACECHECK|Error|Warning" | head
  PMB_MAX_ITER=1 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/profile_step.py kite 1 1 2>&1 | grep -E "RACECHECK|Error|Warning" | head
  PMB_MAX_ITER=3 PMB_PRECOND=1 PMB_FILTER_LS=1 compute-sanitizer --tool memcheck python tools/profile_step.py mobile_robot 8 2 2>&1 | grep -E "ERROR SUMMARY"
  PMB_MAX_ITER=3 compute-sanitizer --tool memcheck python tools/profile_step.py mobile_robot 8 2 2>&1 | grep -E "ERROR SUMMARY"
  PMB_MAX_ITER=1 compute-sanitizer --tool memcheck python tools/profile_step.py kite 2 1 2>&1 | grep -E "ERROR SUMMARY"
done
