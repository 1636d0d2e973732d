#!/bin/bash
# compute-sanitizer over small frames (tools/diag/sanitize_target.py): memcheck, racecheck, synccheck -> gpurun_out/san_<tool>.log
tag=${1:-r2}
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/diag/sanitize_target.py > gpurun_out/san_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|Error|hazard" gpurun_out/san_$tool.log | head -8
done
{ for tool in memcheck racecheck synccheck; do echo "== compute-sanitizer --tool $tool python tools/diag/sanitize_target.py"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target ok|Error|hazard" gpurun_out/san_$tool.log | head -12; done; } > gpurun_out/${tag}_sanitizer.txt
