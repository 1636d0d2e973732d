#!/bin/bash
# compute-sanitizer over one small frame (smoke) + one sort: memcheck, racecheck, synccheck
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/san_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" gpurun_out/san_$tool.log | head -8
done
