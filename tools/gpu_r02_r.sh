#!/bin/bash
# round 2, pass R: certificate restricted to later facets; ncu of the C2 kernel with the certificate
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02r_tests.txt
tail -3 gpurun_out/r02r_tests.txt
B="python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 100000"
$B 2>/dev/null | python tools/bench_line.py "C2"
python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 later-facets certificate"
python bench.py --config c5 --c5-photons 2e8 --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5 n1 2e8', d['ms_per_step'], d['phases_ms'], d['roofline']['trace_only']['frac'])"
ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/r02r_prof_c2 \
    python bench.py --no-cpu --no-e2e --no-api --verify 0 --steps 2 > gpurun_out/r02r_ncu_c2.log 2>&1
ls -la gpurun_out | grep r02r
