#!/bin/bash
# round 2, pass V: culled vs exhaustive search on millions of photons; grid-margin A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -k "exhaustive or certificate or overlap or steep or rowland or config3" 2>&1 | tail -6
python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 hops0"
MXB_GRID_HOPS=2 python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 hops2"
B="python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 100000"
$B 2>/dev/null | python tools/bench_line.py "C2 hops0"
MXB_GRID_HOPS=2 $B 2>/dev/null | python tools/bench_line.py "C2 hops2"
