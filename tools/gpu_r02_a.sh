#!/bin/bash
# round 2, pass A: new parity tests, full GPU suite, both bench arms
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
python -m pytest tests/test_parity_round2.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02a_tests_new.txt
cat gpurun_out/r02a_tests_new.txt | tail -15
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02a_tests_all.txt
tail -5 gpurun_out/r02a_tests_all.txt
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/r02a_bench_ref.json
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
