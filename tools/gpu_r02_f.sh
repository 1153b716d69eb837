#!/bin/bash
# round 2, pass F: previously failing tests, then C3 / C4 bench arms and a lean-e2e chunk sweep
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -k "chandra_c2_golden or rowland_instrument or exported_draws or scalar_process or lean_host" 2>&1 | tail -40 > gpurun_out/r02f_tests.txt
tail -30 gpurun_out/r02f_tests.txt
python bench.py --config c3 --steps 5 > gpurun_out/r02f_c3.json 2> gpurun_out/r02f_c3.err; tail -c 1200 gpurun_out/r02f_c3.json; tail -3 gpurun_out/r02f_c3.err
python bench.py --config c4 --steps 5 > gpurun_out/r02f_c4.json 2> gpurun_out/r02f_c4.err; tail -c 1200 gpurun_out/r02f_c4.json; tail -3 gpurun_out/r02f_c4.err
for ch in 131072 262144 524288 1048576; do
  MXB_HOST_CHUNK=$ch python bench.py --steps 5 --no-cpu --verify 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $ch e2e %.4g lean %.4g'%(d['e2e']['value'], d['e2e']['lean']['value']))"
done
