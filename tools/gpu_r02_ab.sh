#!/bin/bash
# round 2, pass AB: finer culling cells for small programs (A/B on the same box)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -k "chandra or exhaustive or acis or detector or c2 or rowland" 2>&1 | tail -4
B="python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 100000"
for rep in 1 2; do
$B 2>/dev/null | python tools/bench_line.py "C2 fine grid"
MXB_GRID_FINE=0 $B 2>/dev/null | python tools/bench_line.py "C2 coarse grid"
done
