#!/bin/bash
# round 2, pass Z: split kernels with warp-private queues
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -k "config3 or cat or rowland or exhaustive or steep or overlap or stack" 2>&1 | tail -5
python bench.py --config c3 --steps 5 2>gpurun_out/r02z_c3.err | python tools/bench_line.py "C3 split (warp queues)"; tail -2 gpurun_out/r02z_c3.err
MXB_JIT_SPLIT=0 python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 no split"
