#!/bin/bash
# round 2, pass P: statistics kernels after the MLP / parallel-pick changes; lean e2e chunk sweep
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -k "sigma_clip or analysis or capture_res or tolerancing or run_tolerances" 2>&1 | tail -5
python tools/bench_stats.py 1e7 | tee gpurun_out/r02p_stats.json
python tools/bench_stats.py 1e8 | tee -a gpurun_out/r02p_stats.json
for ch in 262144 524288 1048576 2097152; do
  MXB_HOST_CHUNK=$ch python bench.py --steps 20 --no-cpu --no-api --verify 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $ch e2e %.4g lean %.4g'%(d['e2e']['value'], d['e2e']['lean']['value']))"
done
