#!/bin/bash
# round 2, pass AA: own general sincos + shared multilayer bracket (C4 / C5 birth), A/B on the same box
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02aa_tests.txt
tail -4 gpurun_out/r02aa_tests.txt
python bench.py --config c4 --steps 5 2>/dev/null | python tools/bench_line.py "C4 new"
MXB_JIT_DEFINES="-DMXB_LIBM_SINCOS" python bench.py --config c4 --steps 5 2>/dev/null | python tools/bench_line.py "C4 libm_sincos"
python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 new"
python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 100000 2>/dev/null | python tools/bench_line.py "C2 new"
python tools/run_c5.py --photons 1e8 2>/dev/null | tail -1 | cut -c1-400
MXB_JIT_DEFINES="-DMXB_LIBM_SINCOS" python tools/run_c5.py --photons 1e8 2>/dev/null | tail -1 | cut -c1-400
