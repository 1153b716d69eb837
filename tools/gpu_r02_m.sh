#!/bin/bash
# round 2, pass M: search test with the short dependency chain; scan loops without unrolling
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02m_tests.txt
tail -4 gpurun_out/r02m_tests.txt
B="python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 20000"
$B 2>/dev/null | python tools/bench_line.py "C2 new"
MXB_JIT_DEFINES="-DMXB_SEARCH_REFERENCE_ORDER" $B 2>/dev/null | python tools/bench_line.py "C2 search_ref_order"
MXB_JIT_DEFINES="-DMXB_SCAN_UNROLL1" $B 2>/dev/null | python tools/bench_line.py "C2 scan_unroll1"
MXB_JIT_DEFINES="-DMXB_NO_EXPECT" $B 2>/dev/null | python tools/bench_line.py "C2 no_expect"
python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 new"
MXB_JIT_DEFINES="-DMXB_SEARCH_REFERENCE_ORDER" python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 search_ref_order"
MXB_JIT_DEFINES="-DMXB_SCAN_UNROLL1" python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 scan_unroll1"
python bench.py --config c4 --steps 5 2>/dev/null | python tools/bench_line.py "C4 new"
