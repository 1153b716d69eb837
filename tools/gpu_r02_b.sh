#!/bin/bash
# round 2, pass B: re-run the failing tests with tracebacks, C5 on one GPU at reduced size, ncu of the C2 kernel
mkdir -p gpurun_out
python -m pytest tests/test_parity_round2.py -m gpu -q -x --tb=long 2>&1 | tail -80 > gpurun_out/r02b_tests_new.txt
tail -5 gpurun_out/r02b_tests_new.txt
python bench.py --config c5 --c5-photons 2e8 --steps 3 > gpurun_out/r02b_c5_n1.json 2> gpurun_out/r02b_c5_n1.err
tail -c 2500 gpurun_out/r02b_c5_n1.json; tail -5 gpurun_out/r02b_c5_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --verify 0 > gpurun_out/r02b_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/r02b_prof \
    python bench.py --no-cpu --no-e2e --verify 0 --steps 2 > gpurun_out/r02b_ncu.log 2>&1
ls -la gpurun_out/
