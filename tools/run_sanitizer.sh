#!/bin/bash
# compute-sanitizer memcheck / racecheck / synccheck over tools/sanitize_cases.py.
# Usage (on the GPU box): bash tools/run_sanitizer.sh  -> gpurun_out/sanitizer_<tool>.txt
set -u
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py all \
      > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "exit code: $?" >> gpurun_out/sanitizer_$tool.txt
  tail -5 gpurun_out/sanitizer_$tool.txt
done
