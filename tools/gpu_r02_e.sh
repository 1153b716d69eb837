#!/bin/bash
# round 2, pass E: full GPU suite with the tightened tolerances, smoke, compute-sanitizer log
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02e_tests_all.txt
tail -25 gpurun_out/r02e_tests_all.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 bash tools/sanitize.sh > gpurun_out/r02e_sanitizer.log 2>&1; echo "sanitizer rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|== compute|workload ok" gpurun_out/r02e_sanitizer.log
