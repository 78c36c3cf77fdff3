#!/bin/bash
# compute-sanitizer racecheck + memcheck of the fused kernel on a few instances of each problem family (run on a GPU box)
set -x
PMB_MAX_ITER=3 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/profile_step.py mobile_robot 4 1 2>&1 | grep -E "RACECHECK|Error|Warning" | head
PMB_MAX_ITER=2 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/profile_step.py cstr 3 1 2>&1 | grep -E "RACECHECK|Error|Warning" | head
PMB_MAX_ITER=3 compute-sanitizer --tool memcheck python tools/profile_step.py mobile_robot 8 1 2>&1 | grep -E "ERROR SUMMARY"
