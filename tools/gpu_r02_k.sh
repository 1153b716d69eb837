#!/bin/bash
# round 2, pass K: full GPU suite on the new math, benches, ncu captures of the C2 / C3 / C4 kernels as shipped
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02k_tests.txt
tail -6 gpurun_out/r02k_tests.txt
line() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; v=d.get('verify') or {}; print('$1 kernel_ms %.4f frac %.4f verify %s pol %s  %s %s'%(r['kernel_ms'], r['frac'], v.get('indices_bit_exact'), (v.get('max_rel_err') or {}).get('polarization'), d['kernel_path'][:70], {k:(float("%.4g" % x) if isinstance(x,float) else x) for k,x in (d.get('checks') or {}).items() if k in ('on_facet','on_ccd','mean_order','ccd_hit_fraction','on_detector','mean_probability_detected')}))"; }
python bench.py --steps 100 --no-cpu --no-e2e --verify 20000 2>/dev/null | line "C2 new"
MXB_JIT_DEFINES="-DMXB_NO_EXPECT" python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 20000 2>/dev/null | line "C2 no_expect"
MXB_JIT_DEFINES="-DMXB_INLINE_STATUS" python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 20000 2>/dev/null | line "C2 inline_status"
python bench.py --config c3 --steps 5 2>/dev/null | line "C3 new"
python bench.py --config c4 --steps 5 2>/dev/null | line "C4 new"
python bench.py --steps 100 --no-cpu --no-e2e --verify 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('element_api', d.get('element_api'))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02k_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-api --verify 0 > gpurun_out/r02k_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/r02k_prof_c2 \
    python bench.py --no-cpu --no-e2e --no-api --verify 0 --steps 2 > gpurun_out/r02k_ncu_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/r02k_prof_c3 \
    python bench.py --config c3 --photons 9999872 --steps 2 > gpurun_out/r02k_ncu_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/r02k_prof_c4 \
    python bench.py --config c4 --photons 9999872 --steps 2 > gpurun_out/r02k_ncu_c4.log 2>&1
ls -la gpurun_out/ | grep r02k
