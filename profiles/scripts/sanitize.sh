#!/bin/bash
# compute-sanitizer targets for the warp-specialised kernels: memcheck and racecheck over the
# smoke-sized forward (both precision modes; every tcgen05 kernel runs) -- small shapes, the tools are 10-100x slower.
#   bash profiles/scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck]   (default: memcheck then racecheck)
set -u
mkdir -p gpurun_out
tools=${1:-"memcheck racecheck"}
for tool in $tools; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --launch-timeout 120 \
      python __graft_entry__.py --smoke-only > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke\[" gpurun_out/sanitize_$tool.log | tail -5
done
