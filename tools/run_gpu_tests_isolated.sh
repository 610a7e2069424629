#!/bin/bash
# Runs every GPU test in its own process (a faulting kernel cannot poison the others), with a per-test
# timeout; writes gpurun_out/gpu_tests.log and a PASS/FAIL summary.  Usage: tools/run_gpu_tests_isolated.sh [pytest -k expr]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/gpu_tests.log
: > $LOG
rm -f gpurun_out/parity_report.jsonl
IDS=$(python -m pytest tests -m gpu --collect-only -q ${1:+-k "$1"} 2>/dev/null | grep "::")
pass=0; fail=0
for id in $IDS; do
  echo "=== $id" >> $LOG
  timeout 300 python -m pytest "$id" -x -q -p no:cacheprovider >> $LOG 2>&1
  rc=$?
  if [ $rc -eq 0 ]; then pass=$((pass+1)); echo "PASS $id"; else fail=$((fail+1)); echo "FAIL($rc) $id"; fi
done | tee gpurun_out/gpu_tests_summary.txt
echo "done" >> gpurun_out/gpu_tests_summary.txt
grep -c PASS gpurun_out/gpu_tests_summary.txt; grep FAIL gpurun_out/gpu_tests_summary.txt | head -50
