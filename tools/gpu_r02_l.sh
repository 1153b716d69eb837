#!/bin/bash
# round 2, pass L: division-free scan test (+ no unrolling of the scan loops), shared Philox experiment
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02l_tests.txt
tail -4 gpurun_out/r02l_tests.txt
B="python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 20000"
$B 2>/dev/null | python tools/bench_line.py "C2 new"
MXB_JIT_DEFINES="-DMXB_SCAN_EXACT" $B 2>/dev/null | python tools/bench_line.py "C2 scan_exact"
MXB_JIT_DEFINES="-DMXB_SHARE_PHILOX" $B 2>/dev/null | python tools/bench_line.py "C2 share_philox"
MXB_JIT_DEFINES="-DMXB_NO_EXPECT" $B 2>/dev/null | python tools/bench_line.py "C2 no_expect"
MXB_JIT_DEFINES="-DMXB_INLINE_STATUS" $B 2>/dev/null | python tools/bench_line.py "C2 inline_status"
python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 new"
MXB_JIT_DEFINES="-DMXB_SCAN_EXACT" python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 scan_exact"
MXB_JIT_DEFINES="-DMXB_SHARE_PHILOX" python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 share_philox"
python bench.py --config c4 --steps 5 2>/dev/null | python tools/bench_line.py "C4 new"
MXB_JIT_DEFINES="-DMXB_SHARE_PHILOX" python bench.py --config c4 --steps 5 2>/dev/null | python tools/bench_line.py "C4 share_philox"
