#!/bin/bash
# round 2, pass J: fast-build div / sqrt everywhere, own sincos for azimuths, baked scatter widths; C4 regression hunt
mkdir -p gpurun_out
python tools/debug_c4.py 1e8 2>&1 | tail -4
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02j_tests.txt
tail -12 gpurun_out/r02j_tests.txt
line() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; v=d.get('verify') or {}; print('$1 kernel_ms %.4f frac %.4f verify %s pol %s  %s %s'%(r['kernel_ms'], r['frac'], v.get('indices_bit_exact'), (v.get('max_rel_err') or {}).get('polarization'), d['kernel_path'][:70], {k:(round(x,6) if isinstance(x,float) else x) for k,x in (d.get('checks') or {}).items() if k in ('on_facet','on_ccd','mean_order','ccd_hit_fraction','on_detector','mean_probability_detected')}))"; }
python bench.py --steps 100 --no-cpu --no-e2e --verify 20000 2>/dev/null | line "C2 new"
MXB_JIT_THREADS=768 python bench.py --steps 100 --no-cpu --no-e2e --verify 20000 2>/dev/null | line "C2 t768"
MXB_JIT_DEFINES="-DMXB_COLD_STATUS" python bench.py --steps 100 --no-cpu --no-e2e --verify 20000 2>/dev/null | line "C2 cold_status"
python bench.py --config c3 --steps 5 2>/dev/null | line "C3 new"
MXB_JIT_DEFINES="-DMXB_OUTLINE_SCAN" python bench.py --config c3 --steps 5 2>/dev/null | line "C3 outline_scan"
MXB_JIT_DEFINES="-DMXB_COLD_STATUS" python bench.py --config c3 --steps 5 2>/dev/null | line "C3 cold_status"
MXB_JIT_DEFINES="-DMXB_LIBM_SINCOS" python bench.py --config c3 --steps 5 2>/dev/null | line "C3 libm_sincos"
python bench.py --config c4 --steps 5 2>/dev/null | line "C4 new"
python bench.py --config c5 --c5-photons 2e8 --steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5 n1 2e8', d['ms_per_step'], d['phases_ms'], d['roofline']['trace_only'])"
