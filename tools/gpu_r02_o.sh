#!/bin/bash
# round 2, pass O: div2 (shared reciprocal) check, statistics-kernel throughput
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 > gpurun_out/r02o_tests.txt
tail -3 gpurun_out/r02o_tests.txt
B="python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 100000"
$B 2>/dev/null | python tools/bench_line.py "C2 div2"
$B 2>/dev/null | python tools/bench_line.py "C2 div2 again"
python bench.py --config c4 --steps 5 2>/dev/null | python tools/bench_line.py "C4 div2"
python tools/bench_stats.py 1e7 | tee gpurun_out/r02o_stats.json
python tools/bench_stats.py 1e8 | tee -a gpurun_out/r02o_stats.json
