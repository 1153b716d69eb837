#!/bin/bash
# round 2, pass S: successor lists (certificate mode 2)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02s_tests.txt
tail -3 gpurun_out/r02s_tests.txt
B="python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 100000"
$B 2>/dev/null | python tools/bench_line.py "C2"
python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 successor lists"
MXB_JIT_DEFINES="-DMXB_NO_CERT" python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 no_cert"
ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/r02s_prof_c3 \
    python bench.py --config c3 --photons 9999872 --steps 2 > gpurun_out/r02s_ncu_c3.log 2>&1
ls -la gpurun_out | grep r02s
