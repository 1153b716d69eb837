#!/bin/bash
# round 2, pass F (final records, after the last kernel change): records of the tree as committed - tests, smoke, both bench arms, C3 / C4 / C5 arms, launch list,
# ncu captures of the three kernels, sanitizer
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02f_tests.txt
tail -3 gpurun_out/r02f_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02f_bench_reference_n1.json
python bench.py 2>/dev/null | tail -1 > gpurun_out/r02f_bench_n1.json
python tools/bench_line.py "C2 final" < gpurun_out/r02f_bench_n1.json
python bench.py --config c3 2>/dev/null | tail -1 > gpurun_out/r02f_bench_c3.json; python tools/bench_line.py "C3 final" < gpurun_out/r02f_bench_c3.json
python bench.py --config c4 2>/dev/null | tail -1 > gpurun_out/r02f_bench_c4.json; python tools/bench_line.py "C4 final" < gpurun_out/r02f_bench_c4.json
python bench.py --config c5 --c5-photons 2e8 --steps 3 2>/dev/null | tail -1 > gpurun_out/r02f_c5_n1_2e8.json
python -c "import json; d=json.load(open('gpurun_out/r02f_c5_n1_2e8.json')); print('C5 n1 2e8', d['ms_per_step'], d['phases_ms'], d['roofline']['trace_only']['frac'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-api --verify 0 > gpurun_out/r02f_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/r02f_prof_c2 \
    python bench.py --no-cpu --no-e2e --no-api --verify 0 --steps 2 > gpurun_out/r02f_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/r02f_prof_c3 \
    python bench.py --config c3 --photons 9999872 --steps 2 > gpurun_out/r02f_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/r02f_prof_c4 \
    python bench.py --config c4 --photons 9999872 --steps 2 > gpurun_out/r02f_ncu_c4.log 2>&1
bash tools/sanitize.sh > gpurun_out/r02f_sanitizer.log 2>&1; tail -3 gpurun_out/r02f_sanitizer.log
ls gpurun_out | grep r02n
python tools/bench_stats.py 1e7 > gpurun_out/r02f_stats.json; python tools/bench_stats.py 1e8 >> gpurun_out/r02f_stats.json; cat gpurun_out/r02f_stats.json | cut -c1-200
