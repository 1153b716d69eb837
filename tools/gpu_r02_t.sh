#!/bin/bash
# round 2, pass T: interpolation search for global tables, L1 prefetch of the efficiency-table rows, tighter cell lists
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02t_tests.txt
tail -3 gpurun_out/r02t_tests.txt
B="python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 100000"
$B 2>/dev/null | python tools/bench_line.py "C2"
python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3"
MXB_JIT_DEFINES="-DMXB_NO_CERT" python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 no_cert"
python bench.py --config c4 --steps 5 2>/dev/null | python tools/bench_line.py "C4"
