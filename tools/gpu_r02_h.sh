#!/bin/bash
# round 2, pass H: Rodrigues transport / vector rotation / constant-bank tables; phase-locked CTAs (JIT_PHASE)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -25 > gpurun_out/r02h_tests.txt
tail -8 gpurun_out/r02h_tests.txt
line() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; v=d.get('verify') or {}; print('$1 kernel_ms %.4f frac %.4f verify %s pol %s  %s %s'%(r['kernel_ms'], r['frac'], v.get('indices_bit_exact'), (v.get('max_rel_err') or {}).get('polarization'), d['kernel_path'][:70], {k:(round(x,6) if isinstance(x,float) else x) for k,x in (d.get('checks') or {}).items() if k in ('on_facet','on_ccd','mean_order','ccd_hit_fraction','on_detector','mean_probability_detected')}))"; }
python bench.py --steps 100 --no-cpu --no-e2e --verify 20000 2>/dev/null | line "C2 new"
MXB_JIT_DEFINES="-DMXB_PT_FRAME" python bench.py --steps 100 --no-cpu --no-e2e --verify 20000 2>/dev/null | line "C2 pt_frame"
MXB_JIT_DEFINES="-DMXB_ROT_MATRIX" python bench.py --steps 100 --no-cpu --no-e2e --verify 20000 2>/dev/null | line "C2 rot_matrix"
MXB_JIT_DEFINES="-DMXB_NO_CONST_TABLES" python bench.py --steps 100 --no-cpu --no-e2e --verify 20000 2>/dev/null | line "C2 no_const"
MXB_JIT_DEFINES="-DMXB_PT_FRAME -DMXB_ROT_MATRIX -DMXB_NO_CONST_TABLES" python bench.py --steps 100 --no-cpu --no-e2e --verify 20000 2>/dev/null | line "C2 old"
MXB_JIT_PHASE=1 python bench.py --steps 100 --no-cpu --no-e2e --verify 20000 2>/dev/null | line "C2 phase"
python bench.py --config c3 --steps 5 2>/dev/null | line "C3 new"
MXB_JIT_DEFINES="-DMXB_PT_FRAME -DMXB_ROT_MATRIX -DMXB_NO_CONST_TABLES" python bench.py --config c3 --steps 5 2>/dev/null | line "C3 old"
MXB_JIT_PHASE=1 python bench.py --config c3 --steps 5 2>/dev/null | line "C3 phase"
python bench.py --config c4 --steps 5 2>/dev/null | line "C4 new"
MXB_JIT_PHASE=1 python bench.py --config c4 --steps 5 2>/dev/null | line "C4 phase"
