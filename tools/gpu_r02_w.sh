#!/bin/bash
# round 2, pass W: A/B of the interpolation search and of streaming input loads (same box)
mkdir -p gpurun_out
for rep in 1 2; do
python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 default"
MXB_JIT_DEFINES="-DMXB_NO_INTERP_SEARCH" python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 no_interp_search"
MXB_JIT_DEFINES="-DJIT_STREAM_IN" python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 stream_in"
done
B="python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 100000"
$B 2>/dev/null | python tools/bench_line.py "C2 default"
MXB_JIT_DEFINES="-DJIT_STREAM_IN" $B 2>/dev/null | python tools/bench_line.py "C2 stream_in"
