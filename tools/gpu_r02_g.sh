#!/bin/bash
# round 2, pass G: kernel experiments (cold-path outlining, one-instruction prefetch) + ncu capture of C3
mkdir -p gpurun_out
for pf in 1 2; do
  MXB_JIT_PREFETCH=$pf python bench.py --steps 100 --no-cpu --no-e2e --verify 20000 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C2 prefetch=$pf kernel_ms %.4f frac %.4f verify %s %s'%(d['roofline']['kernel_ms'], d['roofline']['frac'], d['verify']['indices_bit_exact'], d['kernel_path'][:60]))"
done
python bench.py --config c3 --steps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3 kernel_ms %.3f frac %.4f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), d['checks'], d['kernel_path'][:80])"
python bench.py --config c4 --steps 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C4 kernel_ms %.3f frac %.4f'%(d['roofline']['kernel_ms'], d['roofline']['frac']), d['kernel_path'][:80])"
ncu --set full --clock-control none --import-source on -k regex:mxb_jit -c 1 -s 3 -f -o gpurun_out/r02g_prof_c3 \
    python bench.py --config c3 --photons 10000000 --steps 2 > gpurun_out/r02g_ncu_c3.log 2>&1
tail -2 gpurun_out/r02g_ncu_c3.log | cut -c1-300
