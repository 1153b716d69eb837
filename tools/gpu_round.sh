#!/bin/bash
# One GPU-box pass: parity tests, both bench arms, the launch list and one full ncu capture of the
# trace kernel.  Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref_$tag.json
python bench.py 2>&1 | tail -1 > gpurun_out/bench_$tag.json
cat gpurun_out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/prof_$tag \
    python bench.py --no-cpu --no-e2e --steps 2 > gpurun_out/ncu_$tag.log 2>&1
