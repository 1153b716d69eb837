#!/bin/bash
# round 2, pass Y: split kernels (regrouping queue between the first facet search and the facet body)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02y_tests.txt
tail -5 gpurun_out/r02y_tests.txt
python bench.py --config c3 --steps 5 2>gpurun_out/r02y_c3.err | python tools/bench_line.py "C3 split"; tail -2 gpurun_out/r02y_c3.err
MXB_JIT_SPLIT=0 python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 no split"
B="python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 100000"
$B 2>/dev/null | python tools/bench_line.py "C2"
