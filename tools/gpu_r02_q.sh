#!/bin/bash
# round 2, pass Q: disjointness certificate (single-hit stop); statistics kernels after the MLP / parallel-pick changes
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 > gpurun_out/r02q_tests.txt
tail -4 gpurun_out/r02q_tests.txt
B="python bench.py --steps 100 --no-cpu --no-e2e --no-api --verify 100000"
$B 2>/dev/null | python tools/bench_line.py "C2 cert"
MXB_JIT_DEFINES="-DMXB_NO_CERT" $B 2>/dev/null | python tools/bench_line.py "C2 no_cert"
python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 cert"
MXB_JIT_DEFINES="-DMXB_NO_CERT" python bench.py --config c3 --steps 5 2>/dev/null | python tools/bench_line.py "C3 no_cert"
python tools/bench_stats.py 1e7 | tee gpurun_out/r02q_stats.json
python tools/bench_stats.py 1e8 | tee -a gpurun_out/r02q_stats.json
for ch in 262144 524288 2097152; do
  MXB_HOST_CHUNK=$ch python bench.py --steps 20 --no-cpu --no-api --verify 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $ch e2e %.4g lean %.4g'%(d['e2e']['value'], d['e2e']['lean']['value']))"
done
