#!/bin/bash
# compute-sanitizer memcheck + racecheck (+ initcheck / synccheck) over every libnsb kernel (race detection).
# Run on a GPU box:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh'   -> gpurun_out/sanitize_*.txt, summary on stdout.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  extra=""
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  SAN_RAYS=${SAN_RAYS:-96} timeout ${SAN_TIMEOUT:-600} $CS --tool $tool $extra --print-limit 20 \
      python tools/sanitize_driver.py > gpurun_out/sanitize_$tool.txt 2>&1
  echo "== $tool: exit $? =="
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_driver: ok|Error:|hazard" gpurun_out/sanitize_$tool.txt | sort | uniq -c | head -12
done
